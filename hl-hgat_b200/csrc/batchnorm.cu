// Training-mode BatchNorm over rows fused with the activation and with the strided store into the
// dense-connection buffer (hl_bn_act_fwd / hl_bn_act_bwd).  Deterministic two-stage column
// reductions: per-block sums of (x - x[0,col]) and its square (the "shifted data" variance algorithm,
// shift = first row), merged in fp64 in a fixed order by one warp per column.  No atomics.
#include "common.cuh"

namespace hl {

constexpr int kBnThreads = 256;
constexpr int kBnWarps = kBnThreads / 32;
constexpr int kBnRowsPerBlock = 128;

static inline int bn_row_blocks(int32_t nrows) { return (nrows + kBnRowsPerBlock - 1) / kBnRowsPerBlock; }

// Thread <-> element mapping shared by the four row-streaming kernels: a warp covers `cw` column chunks of V floats
// (cw = power of two >= min(chunks, 32)) times 32 / cw consecutive rows, so narrow layers (F = 64: 16 chunks) keep all
// 32 lanes busy instead of idling half of them; a thread always works on the SAME columns, so the per-column
// parameters are loaded once, and its rows are walked four at a time with all loads of the four rows in flight.
struct BnMap {
  int cw, rpw, chunk, sub;      // lanes across columns, rows per warp step, this thread's chunk / row slot
  __device__ __forceinline__ BnMap(int chunks_per_group, int lane) {
    cw = 1;
    while (cw < chunks_per_group && cw < 32) cw <<= 1;
    rpw = 32 / cw;
    chunk = lane & (cw - 1);
    sub = lane / cw;
  }
};

constexpr int kBnUnroll = 4;

// partial[(blk * 2 + {0,1}) * width + col] = {sum (x - x[0,col]), sum (x - x[0,col])^2} over the block's rows (fp64)
template <int V>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_partial_kernel(const float* __restrict__ x, int64_t ld_x, int32_t nrows_cap, const int32_t* __restrict__ nvalid,
                        int32_t width, double* __restrict__ partial) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int32_t nrows = nvalid ? min(__ldg(nvalid), nrows_cap) : nrows_cap;
  __shared__ float sh[2][kBnWarps][32 * V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (min(width - (int)blockIdx.y * (32 * V), 32 * V) + V - 1) / V;
  const BnMap mp(chunks, lane);
  const int col = blockIdx.y * (32 * V) + mp.chunk * V;
  const int r0 = blockIdx.x * kBnRowsPerBlock;
  const int r1 = min(r0 + kBnRowsPerBlock, nrows);
  const bool act = mp.chunk < chunks && col < width;
  Pack<V> shift, s1, s2;
#pragma unroll
  for (int i = 0; i < V; ++i) shift.v[i] = s1.v[i] = s2.v[i] = 0.f;
  if (act && r0 < r1) {
    shift = ld_pack<V>(x + col);                             // common shift: row 0 of every column
    const int step = kBnWarps * mp.rpw;
    int r = r0 + warp * mp.rpw + mp.sub;
    for (; r + (kBnUnroll - 1) * step < r1; r += kBnUnroll * step) {
      Pack<V> v[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) v[u] = ld_pack<V>(x + (int64_t)(r + u * step) * ld_x + col);
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float d = v[u].v[i] - shift.v[i];
          s1.v[i] += d;
          s2.v[i] = fmaf(d, d, s2.v[i]);
        }
    }
    for (; r < r1; r += step) {
      Pack<V> v = ld_pack<V>(x + (int64_t)r * ld_x + col);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = v.v[i] - shift.v[i];
        s1.v[i] += d;
        s2.v[i] = fmaf(d, d, s2.v[i]);
      }
    }
  }
  // the row slots of a warp that share a column chunk (lanes chunk, chunk + cw, ...): fixed butterfly order
  for (int o = mp.cw; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1.v[i] += __shfl_xor_sync(0xffffffffu, s1.v[i], o);
      s2.v[i] += __shfl_xor_sync(0xffffffffu, s2.v[i], o);
    }
  if (mp.sub == 0)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sh[0][warp][mp.chunk * V + i] = s1.v[i];
      sh[1][warp][mp.chunk * V + i] = s2.v[i];
    }
  __syncthreads();
  if (warp == 0 && mp.sub == 0 && act) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int w = 0; w < kBnWarps; ++w) {
        a += (double)sh[0][w][mp.chunk * V + i];
        b += (double)sh[1][w][mp.chunk * V + i];
      }
      partial[((int64_t)blockIdx.x * 2 + 0) * width + col + i] = a;
      partial[((int64_t)blockIdx.x * 2 + 1) * width + col + i] = b;
    }
  }
}

// Column sums of the per-row-block partials.  A 256-thread block owns 8 columns: thread (ks, cl) adds the row
// blocks ks, ks + 32, ... of column c0 + cl (64-byte coalesced rows of the partial array), then the 32 slices are
// combined in ascending order through shared memory -- a fixed order, so results are reproducible.  (One warp per
// column walking all row blocks was latency-bound on long batches: 26 us for the 1800 row blocks of a TSP batch.)
constexpr int kBnFinalCols = 8;
__device__ __forceinline__ bool bn_column_sums(const double* __restrict__ partial, int nblk, int32_t width,
                                               double& a, double& b, int& c) {
  __shared__ double sh[2][32][kBnFinalCols];
  const int cl = threadIdx.x & (kBnFinalCols - 1), ks = threadIdx.x / kBnFinalCols;
  c = blockIdx.x * kBnFinalCols + cl;
  a = 0.0; b = 0.0;
  if (c < width)
    for (int k = ks; k < nblk; k += 32) {
      a += partial[((int64_t)k * 2 + 0) * width + c];
      b += partial[((int64_t)k * 2 + 1) * width + c];
    }
  sh[0][ks][cl] = a;
  sh[1][ks][cl] = b;
  __syncthreads();
  if (ks != 0 || c >= width) return false;
  a = 0.0; b = 0.0;
  for (int j = 0; j < 32; ++j) { a += sh[0][j][cl]; b += sh[1][j][cl]; }
  return true;
}

__global__ void bn_stats_final_kernel(const double* __restrict__ partial, int nblk, const float* __restrict__ x,
                                      int32_t nrows_cap, const int32_t* __restrict__ nvalid, int32_t width,
                                      float* __restrict__ stats, float* __restrict__ running_mean,
                                      float* __restrict__ running_var, float momentum,
                                      long long* __restrict__ batches_tracked) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  if (batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *batches_tracked += 1;   // nn.BatchNorm1d.num_batches_tracked
  const int32_t nrows = nvalid ? min(__ldg(nvalid), nrows_cap) : nrows_cap;
  double a, b;
  int c;
  if (bn_column_sums(partial, nblk, width, a, b, c)) {
    if (nrows > 0) {
      const double n = (double)nrows, m = a / n;
      const float mean = (float)((double)__ldg(x + c) + m);
      const float var = (float)fmax(b / n - m * m, 0.0);
      stats[c] = mean;
      stats[width + c] = var;
      if (running_mean) {                                     // nn.BatchNorm1d bookkeeping: unbiased variance
        const float unbias = (float)(n / fmax(n - 1.0, 1.0));
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
      }
    } else {
      stats[c] = 0.f;
      stats[width + c] = 0.f;
    }
  }
}

// Statistics from the per-32-row-block (mean | M2) pairs the producing GEMM's epilogue wrote (hl_gemm2_bn_tf32x3), merged
// in fp64 in two passes of independent loads (no serial Chan chain: a 50,000-row activation has 1,563 blocks):
//   mean = sum_b n_b mean_b / n,      M2 = sum_b (M2_b + n_b (mean_b - mean)^2)
// 1024 threads = 4 columns x 256 block slices; slice sums are combined in a fixed order (shuffle over the 8 slices of a
// warp, then the 32 warps serially): reproducible.
constexpr int kBnTileCols = 4, kBnTileSlices = 256, kBnTileHold = 8;
__device__ __forceinline__ double bn_tiles_reduce(double v, double (*sh)[kBnTileCols], int cl) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  __syncthreads();                                        // the previous use of sh is over
  if ((threadIdx.x & 31) < kBnTileCols) sh[threadIdx.x >> 5][cl] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll 8
  for (int w = 0; w < kBnTileSlices / 8; ++w) t += sh[w][cl];
  return t;
}

__global__ void __launch_bounds__(kBnTileCols* kBnTileSlices)
bn_stats_final_tiles_kernel(const float* __restrict__ part, int32_t nrows_cap, const int32_t* __restrict__ nvalid,
                            int32_t width, float* __restrict__ stats, float* __restrict__ running_mean,
                            float* __restrict__ running_var, float momentum, long long* __restrict__ batches_tracked) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  if (batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *batches_tracked += 1;
  __shared__ double sh[kBnTileSlices / 8][kBnTileCols];
  // every load of the launch is issued up front (the row count and up to kBnTileHold (mean, M2) pairs per thread: enough
  // for 65,536 rows); blocks past *nvalid hold zeros and get weight 0, so the loads do not wait for the row count
  const int nblk = (nrows_cap + 31) / 32;
  const int cl = threadIdx.x & (kBnTileCols - 1), ks = threadIdx.x / kBnTileCols;
  const int c = min(blockIdx.x * kBnTileCols + cl, width - 1);       // overhanging columns: duplicate work, no store
  const bool store = blockIdx.x * kBnTileCols + cl < width && ks == 0;
  const float* p = part + c;
  float mk[kBnTileHold], qk[kBnTileHold];
#pragma unroll
  for (int u = 0; u < kBnTileHold; ++u) {
    const int k = ks + u * kBnTileSlices;
    mk[u] = k < nblk ? __ldg(p + (int64_t)k * 2 * width) : 0.f;
    qk[u] = k < nblk ? __ldg(p + ((int64_t)k * 2 + 1) * width) : 0.f;
  }
  const int32_t nrows = nvalid ? min(__ldg(nvalid), nrows_cap) : nrows_cap;
  double a = 0.0;
#pragma unroll
  for (int u = 0; u < kBnTileHold; ++u)
    a += (double)max(0, min(32, nrows - (ks + u * kBnTileSlices) * 32)) * (double)mk[u];
  for (int k = ks + kBnTileHold * kBnTileSlices; k < nblk; k += kBnTileSlices)
    a += (double)max(0, min(32, nrows - k * 32)) * (double)__ldg(p + (int64_t)k * 2 * width);
  const double n = (double)nrows;
  const double mean = nrows > 0 ? bn_tiles_reduce(a, sh, cl) / n : 0.0;
  double q = 0.0;
#pragma unroll
  for (int u = 0; u < kBnTileHold; ++u) {
    const double d = (double)mk[u] - mean;
    q += (double)qk[u] + (double)max(0, min(32, nrows - (ks + u * kBnTileSlices) * 32)) * d * d;
  }
  for (int k = ks + kBnTileHold * kBnTileSlices; k < nblk; k += kBnTileSlices) {
    const double d = (double)__ldg(p + (int64_t)k * 2 * width) - mean;
    q += (double)__ldg(p + ((int64_t)k * 2 + 1) * width) + (double)max(0, min(32, nrows - k * 32)) * d * d;
  }
  const double m2 = bn_tiles_reduce(q, sh, cl);
  if (!store) return;
  if (nrows > 0) {
    const float fm = (float)mean, var = (float)fmax(m2 / n, 0.0);
    stats[c] = fm;
    stats[width + c] = var;
    if (running_mean) {
      const float unbias = (float)(n / fmax(n - 1.0, 1.0));
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * fm;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  } else {
    stats[c] = 0.f;
    stats[width + c] = 0.f;
  }
}

// y = act(gamma (x - mean) rstd + beta): grid (row blocks of 128, column groups of 32 V), the mapping of BnMap
template <int V>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, int64_t ld_x, int32_t nrows, const int32_t* __restrict__ nvalid, int32_t width,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
                float eps, float slope, float* __restrict__ y, int64_t ld_y) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int32_t nv = nvalid ? min(__ldg(nvalid), nrows) : nrows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (min(width - (int)blockIdx.y * (32 * V), 32 * V) + V - 1) / V;
  const BnMap mp(chunks, lane);
  const int col = blockIdx.y * (32 * V) + mp.chunk * V;
  if (mp.chunk >= chunks || col >= width) return;
  const int r0 = blockIdx.x * kBnRowsPerBlock, r1 = min(r0 + kBnRowsPerBlock, nrows);
  float mean[V], rstd[V], g[V], b[V];                         // this thread's columns: loaded once
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mean[i] = __ldg(stats + col + i);
    rstd[i] = rsqrtf(__ldg(stats + width + col + i) + eps);
    g[i] = gamma ? __ldg(gamma + col + i) : 1.f;
    b[i] = beta ? __ldg(beta + col + i) : 0.f;
  }
  const int step = kBnWarps * mp.rpw;
  for (int r = r0 + warp * mp.rpw + mp.sub; r < r1; r += kBnUnroll * step) {
    Pack<V> v[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int rr = r + u * step;
      if (rr < r1 && rr < nv) v[u] = ld_pack<V>(x + (int64_t)rr * ld_x + col);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int rr = r + u * step;
      if (rr >= r1) break;
      Pack<V> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float t = rr < nv ? (v[u].v[i] - mean[i]) * rstd[i] * g[i] + b[i] : 0.f;
        o.v[i] = t > 0.f ? t : t * slope;                      // ghost (padding) rows stay exactly zero
      }
      st_pack<V>(y + (int64_t)rr * ld_y + col, o);
    }
  }
}

// backward stage 1: partial column sums of dz and dz * xhat (dz = dy * act'(y))
template <int V>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_partial_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ y, int64_t ld_y,
                      const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ dy2, int64_t ld_dy2,
                      int32_t nrows_cap, const int32_t* __restrict__ nvalid,
                      int32_t width, const float* __restrict__ stats, float eps, float slope, double* __restrict__ partial) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  __shared__ float sh[2][kBnWarps][32 * V];
  const int32_t nrows = nvalid ? min(__ldg(nvalid), nrows_cap) : nrows_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (min(width - (int)blockIdx.y * (32 * V), 32 * V) + V - 1) / V;
  const BnMap mp(chunks, lane);
  const int col = blockIdx.y * (32 * V) + mp.chunk * V;
  const int r0 = blockIdx.x * kBnRowsPerBlock;
  const int r1 = min(r0 + kBnRowsPerBlock, nrows);
  const bool act = mp.chunk < chunks && col < width;
  Pack<V> s1, s2;
#pragma unroll
  for (int i = 0; i < V; ++i) s1.v[i] = s2.v[i] = 0.f;
  if (act) {
    float mean[V], rstd[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      mean[i] = __ldg(stats + col + i);
      rstd[i] = rsqrtf(__ldg(stats + width + col + i) + eps);
    }
    const int step = kBnWarps * mp.rpw;
    for (int r = r0 + warp * mp.rpw + mp.sub; r < r1; r += kBnUnroll * step) {
      Pack<V> xv[kBnUnroll], yv[kBnUnroll], gv[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int rr = r + u * step;
        if (rr < r1) {
          xv[u] = ld_pack<V>(x + (int64_t)rr * ld_x + col);
          yv[u] = ld_pack<V>(y + (int64_t)rr * ld_y + col);
          gv[u] = ld_pack<V>(dy + (int64_t)rr * ld_dy + col);
          if (dy2) {                                             // the gradient arrives in two pieces: summed on the fly
            const Pack<V> g2 = ld_pack<V>(dy2 + (int64_t)rr * ld_dy2 + col);
#pragma unroll
            for (int i = 0; i < V; ++i) gv[u].v[i] += g2.v[i];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        if (r + u * step >= r1) break;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float dz = yv[u].v[i] > 0.f ? gv[u].v[i] : gv[u].v[i] * slope;
          s1.v[i] += dz;
          s2.v[i] = fmaf(dz, (xv[u].v[i] - mean[i]) * rstd[i], s2.v[i]);
        }
      }
    }
  }
  for (int o = mp.cw; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1.v[i] += __shfl_xor_sync(0xffffffffu, s1.v[i], o);
      s2.v[i] += __shfl_xor_sync(0xffffffffu, s2.v[i], o);
    }
  if (mp.sub == 0)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sh[0][warp][mp.chunk * V + i] = s1.v[i];
      sh[1][warp][mp.chunk * V + i] = s2.v[i];
    }
  __syncthreads();
  if (warp == 0 && mp.sub == 0 && act) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int w = 0; w < kBnWarps; ++w) {
        a += (double)sh[0][w][mp.chunk * V + i];
        b += (double)sh[1][w][mp.chunk * V + i];
      }
      partial[((int64_t)blockIdx.x * 2 + 0) * width + col + i] = a;
      partial[((int64_t)blockIdx.x * 2 + 1) * width + col + i] = b;
    }
  }
}

// sums[0:F] = sum dz (= dbeta), sums[F:2F] = sum dz*xhat (= dgamma); one warp per column
__global__ void bn_bwd_final_kernel(const double* __restrict__ partial, int nblk, int32_t width,
                                    float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int accumulate) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  double a, b;
  int c;
  if (bn_column_sums(partial, nblk, width, a, b, c)) {
    sums[c] = (float)a;
    sums[width + c] = (float)b;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)a : (float)a;
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)b : (float)b;
  }
}

template <int V>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ y, int64_t ld_y,
                    const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ dy2, int64_t ld_dy2,
                    int32_t nrows, const int32_t* __restrict__ nvalid, int32_t width,
                    const float* __restrict__ gamma, const float* __restrict__ stats, const float* __restrict__ sums,
                    float eps, float slope, float* __restrict__ dx, int64_t ld_dx, const float* __restrict__ inv_count) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int32_t nv = nvalid ? min(__ldg(nvalid), nrows) : nrows;
  const float inv_n = inv_count ? __ldg(inv_count) : 1.f / (float)max(nv, 1);   // inv_count: 1 / rows over ALL ranks (SyncBN)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (min(width - (int)blockIdx.y * (32 * V), 32 * V) + V - 1) / V;
  const BnMap mp(chunks, lane);
  const int col = blockIdx.y * (32 * V) + mp.chunk * V;
  if (mp.chunk >= chunks || col >= width) return;
  const int r0 = blockIdx.x * kBnRowsPerBlock, r1 = min(r0 + kBnRowsPerBlock, nrows);
  float mean[V], rstd[V], g[V], m1[V], m2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mean[i] = __ldg(stats + col + i);
    rstd[i] = rsqrtf(__ldg(stats + width + col + i) + eps);
    g[i] = gamma ? __ldg(gamma + col + i) : 1.f;
    m1[i] = __ldg(sums + col + i) * inv_n;
    m2[i] = __ldg(sums + width + col + i) * inv_n;
  }
  const int step = kBnWarps * mp.rpw;
  for (int r = r0 + warp * mp.rpw + mp.sub; r < r1; r += kBnUnroll * step) {
    Pack<V> xv[kBnUnroll], yv[kBnUnroll], gv[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int rr = r + u * step;
      if (rr < r1 && rr < nv) {
        xv[u] = ld_pack<V>(x + (int64_t)rr * ld_x + col);
        yv[u] = ld_pack<V>(y + (int64_t)rr * ld_y + col);
        gv[u] = ld_pack<V>(dy + (int64_t)rr * ld_dy + col);
        if (dy2) {
          const Pack<V> g2 = ld_pack<V>(dy2 + (int64_t)rr * ld_dy2 + col);
#pragma unroll
          for (int i = 0; i < V; ++i) gv[u].v[i] += g2.v[i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int rr = r + u * step;
      if (rr >= r1) break;
      Pack<V> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (rr < nv) {
          const float dz = yv[u].v[i] > 0.f ? gv[u].v[i] : gv[u].v[i] * slope;
          const float xh = (xv[u].v[i] - mean[i]) * rstd[i];
          o.v[i] = g[i] * rstd[i] * (dz - m1[i] - xh * m2[i]);
        } else {
          o.v[i] = 0.f;
        }
      }
      st_pack<V>(dx + (int64_t)rr * ld_dx + col, o);
    }
  }
}

}  // namespace hl

extern "C" size_t hl_bn_workspace(int32_t nrows, int32_t width) {
  if (nrows < 0 || width < 0) return 0;
  return hl::align_up((size_t)hl::bn_row_blocks(nrows > 0 ? nrows : 1) * 2 * (size_t)width * sizeof(double), 256) +
         hl::align_up(2 * (size_t)width * sizeof(float), 256);
}

extern "C" int hl_bn_act_fwd(const float* x, int64_t ld_x, int32_t nrows, int32_t width,
                             const float* gamma, const float* beta, float eps, float slope,
                             float* y, int64_t ld_y, float* stats, const int32_t* nvalid,
                             float* running_mean, float* running_var, float momentum,
                             int64_t* num_batches_tracked,
                             void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !stats) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_bn_workspace(nrows, width)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  double* partial = reinterpret_cast<double*>(workspace);
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  const int nblk = bn_row_blocks(nrows);
  dim3 grid(nblk, (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_stats_partial_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  else if (V == 2) hl::launch_pdl(bn_stats_partial_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  else hl::launch_pdl(bn_stats_partial_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  HL_LAUNCH_CHECK("bn_stats_partial_kernel");
  hl::launch_pdl(bn_stats_final_kernel, (width + kBnFinalCols - 1) / kBnFinalCols, 256, 0, st, partial, nblk, x, nrows, nvalid, width, stats,
                                                                    running_mean, running_mean ? running_var : nullptr, momentum,
                                                                    reinterpret_cast<long long*>(num_batches_tracked));
  HL_LAUNCH_CHECK("bn_stats_final_kernel");
  if (V == 4) hl::launch_pdl(bn_apply_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else if (V == 2) hl::launch_pdl(bn_apply_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else hl::launch_pdl(bn_apply_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  HL_LAUNCH_CHECK("bn_apply_kernel");
  return HL_OK;
}

// hl_bn_act_fwd with the statistics pass replaced by the block statistics the producing GEMM wrote (`bn_part`,
// [ceil(nrows/32)][2][width]): finalize + apply, two launches instead of three and one pass over x less.
extern "C" int hl_bn_act_fwd_tiles(const float* x, int64_t ld_x, int32_t nrows, int32_t width,
                                   const float* gamma, const float* beta, float eps, float slope,
                                   float* y, int64_t ld_y, float* stats, const int32_t* nvalid,
                                   float* running_mean, float* running_var, float momentum,
                                   int64_t* num_batches_tracked, const float* bn_part, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !stats || !bn_part) return HL_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  hl::launch_pdl(bn_stats_final_tiles_kernel, (width + kBnTileCols - 1) / kBnTileCols, kBnTileCols * kBnTileSlices, 0, st, bn_part, nrows, nvalid, width, stats, running_mean, running_mean ? running_var : nullptr, momentum,
      reinterpret_cast<long long*>(num_batches_tracked));
  HL_LAUNCH_CHECK("bn_stats_final_tiles_kernel");
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  dim3 grid(bn_row_blocks(nrows), (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_apply_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else if (V == 2) hl::launch_pdl(bn_apply_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else hl::launch_pdl(bn_apply_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  HL_LAUNCH_CHECK("bn_apply_kernel");
  return HL_OK;
}

extern "C" int hl_bn_act_bwd(const float* x, int64_t ld_x, const float* y, int64_t ld_y,
                             const float* dy, int64_t ld_dy, const float* dy2, int64_t ld_dy2, int32_t nrows, int32_t width,
                             const float* gamma, const float* stats, float eps, float slope,
                             float* dx, int64_t ld_dx, float* dgamma, float* dbeta, int accumulate_param_grads,
                             const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !dy || !dx || !stats) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_bn_workspace(nrows, width)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int nblk = bn_row_blocks(nrows);
  double* partial = reinterpret_cast<double*>(workspace);
  float* sums = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) +
                                         align_up((size_t)nblk * 2 * (size_t)width * sizeof(double), 256));
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  V = min(V, vec_for(dy, ld_dy, width, V));
  V = min(V, vec_for(dy2, ld_dy2, width, V));
  V = min(V, vec_for(dx, ld_dx, width, V));
  dim3 grid(nblk, (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_bwd_partial_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  else if (V == 2) hl::launch_pdl(bn_bwd_partial_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  else hl::launch_pdl(bn_bwd_partial_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  HL_LAUNCH_CHECK("bn_bwd_partial_kernel");
  hl::launch_pdl(bn_bwd_final_kernel, (width + kBnFinalCols - 1) / kBnFinalCols, 256, 0, st, partial, nblk, width, sums, dgamma, dbeta, accumulate_param_grads);
  HL_LAUNCH_CHECK("bn_bwd_final_kernel");
  if (V == 4) hl::launch_pdl(bn_bwd_apply_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, nullptr);
  else if (V == 2) hl::launch_pdl(bn_bwd_apply_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, nullptr);
  else hl::launch_pdl(bn_bwd_apply_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, nullptr);
  HL_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return HL_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The same kernels as separate phases, for BatchNorm statistics synchronised across data-parallel ranks: the caller
// exchanges the per-rank statistics (forward) / column sums (backward) between the reduction and the apply phase.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int hl_bn_stats(const float* x, int64_t ld_x, int32_t nrows, int32_t width, float* stats,
                           const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !stats) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_bn_workspace(nrows, width)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  double* partial = reinterpret_cast<double*>(workspace);
  const int V = vec_for(x, ld_x, width, 4);
  const int nblk = bn_row_blocks(nrows);
  dim3 grid(nblk, (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_stats_partial_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  else if (V == 2) hl::launch_pdl(bn_stats_partial_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  else hl::launch_pdl(bn_stats_partial_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, partial);
  HL_LAUNCH_CHECK("bn_stats_partial_kernel");
  hl::launch_pdl(bn_stats_final_kernel, (width + kBnFinalCols - 1) / kBnFinalCols, 256, 0, st, partial, nblk, x, nrows, nvalid, width, stats,
                                                                    nullptr, nullptr, 0.f, nullptr);
  HL_LAUNCH_CHECK("bn_stats_final_kernel");
  return HL_OK;
}

extern "C" int hl_bn_apply(const float* x, int64_t ld_x, int32_t nrows, int32_t width, const float* gamma,
                           const float* beta, const float* stats, float eps, float slope, float* y, int64_t ld_y,
                           const int32_t* nvalid, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !stats) return HL_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  dim3 grid(bn_row_blocks(nrows), (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_apply_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else if (V == 2) hl::launch_pdl(bn_apply_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  else hl::launch_pdl(bn_apply_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, nrows, nvalid, width, gamma, beta, stats, eps, slope, y, ld_y);
  HL_LAUNCH_CHECK("bn_apply_kernel");
  return HL_OK;
}

extern "C" int hl_bn_bwd_sums(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy, int64_t ld_dy,
                              int32_t nrows, int32_t width, const float* stats, float eps, float slope, float* sums,
                              const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !dy || !stats || !sums) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_bn_workspace(nrows, width)) return HL_ERR_WORKSPACE;
  const float* dy2 = nullptr;
  const int64_t ld_dy2 = 0;
  cudaStream_t st = as_stream(stream);
  const int nblk = bn_row_blocks(nrows);
  double* partial = reinterpret_cast<double*>(workspace);
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  V = min(V, vec_for(dy, ld_dy, width, V));
  dim3 grid(nblk, (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_bwd_partial_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  else if (V == 2) hl::launch_pdl(bn_bwd_partial_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  else hl::launch_pdl(bn_bwd_partial_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, stats, eps, slope, partial);
  HL_LAUNCH_CHECK("bn_bwd_partial_kernel");
  hl::launch_pdl(bn_bwd_final_kernel, (width + kBnFinalCols - 1) / kBnFinalCols, 256, 0, st, partial, nblk, width, sums, nullptr, nullptr, 0);
  HL_LAUNCH_CHECK("bn_bwd_final_kernel");
  return HL_OK;
}

extern "C" int hl_bn_bwd_apply(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy, int64_t ld_dy,
                               int32_t nrows, int32_t width, const float* gamma, const float* stats, const float* sums,
                               const float* inv_count, float eps, float slope, float* dx, int64_t ld_dx,
                               const int32_t* nvalid, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 1 || width < 1 || !x || !y || !dy || !dx || !stats || !sums) return HL_ERR_INVALID;
  const float* dy2 = nullptr;
  const int64_t ld_dy2 = 0;
  cudaStream_t st = as_stream(stream);
  int V = vec_for(x, ld_x, width, 4);
  V = min(V, vec_for(y, ld_y, width, V));
  V = min(V, vec_for(dy, ld_dy, width, V));
  V = min(V, vec_for(dx, ld_dx, width, V));
  dim3 grid(bn_row_blocks(nrows), (width + 32 * V - 1) / (32 * V));
  if (V == 4) hl::launch_pdl(bn_bwd_apply_kernel<4>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, inv_count);
  else if (V == 2) hl::launch_pdl(bn_bwd_apply_kernel<2>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, inv_count);
  else hl::launch_pdl(bn_bwd_apply_kernel<1>, grid, kBnThreads, 0, st, x, ld_x, y, ld_y, dy, ld_dy, dy2, ld_dy2, nrows, nvalid, width, gamma, stats, sums, eps, slope, dx, ld_dx, inv_count);
  HL_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return HL_OK;
}
