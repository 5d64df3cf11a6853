// Weight / bias gradients of the dense layers as deterministic row reductions (hl_wgrad, hl_colsum).
//
// dW[Fo,Fi] = g[R,Fo]^T x[R,Fi] contracts over the R (= tens of thousands of) rows of the mini-batch while
// the output is tiny (64x28 ... 256x704).  A library GEMM launches one CTA per 64x64 output tile and
// walks all R rows serially (~115 us per call, 60 calls per step in the ZINC model, profiles/r1_*).
// Here the rows are split over S CTAs per output tile (split-K), partial tiles go to a workspace and
// are summed in a fixed order by a second kernel -- no atomics, bit-reproducible.
#include "common.cuh"

namespace hl {

constexpr int kWgTile = 64;       // output tile (Fo x Fi) per CTA
constexpr int kWgRows = 16;       // rows staged per step
constexpr int kWgThreads = 256;   // 16 x 16 threads, 4 x 4 outputs each

// partial[s][fo][fi]
__global__ void __launch_bounds__(kWgThreads)
wgrad_partial_kernel(const float* __restrict__ g, int64_t ld_g, const float* __restrict__ x, int64_t ld_x,
                     int32_t nrows, int32_t fo, int32_t fi, int32_t rows_per_split, float* __restrict__ partial) {
  __shared__ __align__(16) float gs[kWgRows][kWgTile];
  __shared__ __align__(16) float xs[kWgRows][kWgTile];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int fi0 = blockIdx.x * kWgTile, fo0 = blockIdx.y * kWgTile;
  const int r_begin = blockIdx.z * rows_per_split;
  const int r_end = min(r_begin + rows_per_split, nrows);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: thread -> (row = tid / 16, 4 consecutive columns = (tid % 16) * 4)
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
  const bool g_vec = (ld_g % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) && (fo0 + lc + 3 < fo);
  const bool x_vec = (ld_x % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (fi0 + lc + 3 < fi);

  for (int r0 = r_begin; r0 < r_end; r0 += kWgRows) {
    const int r = r0 + lr;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), xv = gv;
    if (r < r_end) {
      const float* gp = g + (int64_t)r * ld_g + fo0 + lc;
      const float* xp = x + (int64_t)r * ld_x + fi0 + lc;
      if (g_vec) gv = __ldg(reinterpret_cast<const float4*>(gp));
      else {
        if (fo0 + lc + 0 < fo) gv.x = __ldg(gp + 0);
        if (fo0 + lc + 1 < fo) gv.y = __ldg(gp + 1);
        if (fo0 + lc + 2 < fo) gv.z = __ldg(gp + 2);
        if (fo0 + lc + 3 < fo) gv.w = __ldg(gp + 3);
      }
      if (x_vec) xv = __ldg(reinterpret_cast<const float4*>(xp));
      else {
        if (fi0 + lc + 0 < fi) xv.x = __ldg(xp + 0);
        if (fi0 + lc + 1 < fi) xv.y = __ldg(xp + 1);
        if (fi0 + lc + 2 < fi) xv.z = __ldg(xp + 2);
        if (fi0 + lc + 3 < fi) xv.w = __ldg(xp + 3);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&gs[lr][lc]) = gv;
    *reinterpret_cast<float4*>(&xs[lr][lc]) = xv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kWgRows; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&gs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&xs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* out = partial + (int64_t)blockIdx.z * fo * fi;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = fo0 + ty * 4 + i;
    if (o >= fo) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = fi0 + tx * 4 + j;
      if (c < fi) out[(int64_t)o * fi + c] = acc[i][j];
    }
  }
}

// dw[i] (=|+=) sum_k partial[k][i]: 8 lanes share one output (k = lane, lane + 8, ... each in ascending order),
// then a fixed shuffle tree combines the 8 sub-sums -- deterministic, and 8x shorter dependent-load chains than
// one thread walking all partials (the serial version was latency-bound: ~10 us for a 64 x 64 gradient).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int32_t splits, int64_t n,
                    float* __restrict__ dw, int64_t ld_dw, int32_t fi, int accumulate) {
  const int sub = threadIdx.x & 7;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < ((n + 31) & ~(int64_t)31);
       i += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    float s = 0.f;
    if (i < n) {
      int k = sub;
      for (; k + 24 < splits; k += 32) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(partial + (int64_t)(k + 8 * u) * n + i);
#pragma unroll
        for (int u = 0; u < 4; ++u) s += v[u];
      }
      for (; k < splits; k += 8) s += __ldg(partial + (int64_t)k * n + i);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0 && i < n) {
      const int64_t o = i / fi, c = i - o * fi;
      float* p = dw + o * ld_dw + c;
      *p = accumulate ? *p + s : s;
    }
  }
}

// column sums: partial[s][f] then fixed-order reduce
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ g, int64_t ld_g, int32_t nrows, int32_t width, int32_t rows_per_split,
                      float* __restrict__ partial) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r_begin = blockIdx.y * rows_per_split, r_end = min(r_begin + rows_per_split, nrows);
  float s = 0.f;
  if (col < width)
    for (int r = r_begin + warp; r < r_end; r += 8) s += __ldg(g + (int64_t)r * ld_g + col);
  sh[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && col < width) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][lane];
    partial[(int64_t)blockIdx.y * width + col] = t;
  }
}

static int pick_splits(int32_t nrows, int tiles) {
  int s = (3 * 148 + tiles - 1) / tiles;                    // ~3 CTAs per SM in total
  const int max_s = (nrows + 4 * kWgRows - 1) / (4 * kWgRows);
  if (s > max_s) s = max_s;
  if (s > 128) s = 128;
  return s < 1 ? 1 : s;
}

}  // namespace hl

extern "C" size_t hl_wgrad_workspace(int32_t nrows, int32_t fo, int32_t fi) {
  if (nrows < 0 || fo < 1 || fi < 1) return 0;
  const int tiles = ((fo + hl::kWgTile - 1) / hl::kWgTile) * ((fi + hl::kWgTile - 1) / hl::kWgTile);
  return (size_t)hl::pick_splits(nrows, tiles) * (size_t)fo * (size_t)fi * sizeof(float) + 256;
}

extern "C" int hl_wgrad(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo, int32_t fi,
                        float* dw, int64_t ld_dw, int accumulate, void* workspace, size_t workspace_bytes,
                        hl_stream_t stream) {
  using namespace hl;
  if (nrows < 0 || fo < 1 || fi < 1 || !dw) return HL_ERR_INVALID;
  if (nrows > 0 && (!g || !x)) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_wgrad_workspace(nrows, fo, fi)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int tx = (fi + kWgTile - 1) / kWgTile, ty = (fo + kWgTile - 1) / kWgTile;
  const int splits = pick_splits(nrows, tx * ty);
  const int rows_per_split = ((nrows + splits - 1) / splits + kWgRows - 1) / kWgRows * kWgRows;
  float* partial = reinterpret_cast<float*>(workspace);
  wgrad_partial_kernel<<<dim3(tx, ty, splits), kWgThreads, 0, st>>>(g, ld_g, x, ld_x, nrows, fo, fi,
                                                                     rows_per_split > 0 ? rows_per_split : kWgRows, partial);
  HL_LAUNCH_CHECK("wgrad_partial_kernel");
  const int64_t n = (int64_t)fo * fi;
  wgrad_reduce_kernel<<<(int)((n * 8 + 255) / 256), 256, 0, st>>>(partial, splits, n, dw, ld_dw, fi, accumulate);
  HL_LAUNCH_CHECK("wgrad_reduce_kernel");
  return HL_OK;
}

extern "C" size_t hl_colsum_workspace(int32_t nrows, int32_t width) {
  if (nrows < 0 || width < 1) return 0;
  return (size_t)128 * (size_t)width * sizeof(float) + 256;
}

extern "C" int hl_colsum(const float* g, int64_t ld_g, int32_t nrows, int32_t width, float* out, int accumulate,
                         void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 0 || width < 1 || !out || (nrows > 0 && !g)) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_colsum_workspace(nrows, width)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int ctiles = (width + 31) / 32;
  int splits = (2 * 148 + ctiles - 1) / ctiles;
  const int max_s = (nrows + 63) / 64;
  if (splits > max_s) splits = max_s;
  if (splits > 128) splits = 128;
  if (splits < 1) splits = 1;
  const int rows_per_split = (nrows + splits - 1) / splits;
  float* partial = reinterpret_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3(ctiles, splits), 256, 0, st>>>(g, ld_g, nrows, width, rows_per_split > 0 ? rows_per_split : 1, partial);
  HL_LAUNCH_CHECK("colsum_partial_kernel");
  wgrad_reduce_kernel<<<(width * 8 + 255) / 256, 256, 0, st>>>(partial, splits, width, out, width, width, accumulate);
  HL_LAUNCH_CHECK("wgrad_reduce_kernel");
  return HL_OK;
}
