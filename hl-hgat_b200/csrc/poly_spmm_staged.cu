// Row-window staged polynomial SpMM for short-row, banded operators (block-diagonal mini-batches of
// small graphs: ZINC / peptides L0 and L1, 3-4 nnz per row, neighbours a few rows away).
//
// The per-row kernel in poly_spmm.cu is latency-bound there: every row is a chain of dependent global
// loads (rowptr -> colidx/vals -> x rows) with only a few hundred bytes behind it (ncu: long-scoreboard
// stalls, 31% DRAM).  Here a persistent CTA walks 48 KB row tiles (192 rows at F = 64) through a 3-stage
// shared-memory ring:
//   * one producer warp streams the tile's x rows, its rowptr slice and its colidx / vals slices into
//     shared memory with cp.async.bulk (TMA engine, completion counted on an mbarrier), two tiles ahead
//     of the math and with the next tile's nnz range prefetched, so HBM sees large coalesced requests
//     with ~150 KB in flight per SM and no thread ever waits on a dependent global load chain;
//   * 16 consumer warps (F/8 lanes x 2 float4 per row) read indices and source rows from
//     shared memory (neighbours outside the tile's own row window fall back to an L2/global gather),
//     run the same ordered multiply/add chain and recurrence epilogue as the per-row kernel -- results
//     are bit-identical -- and write the output rows with 128-bit stores.
#include "common.cuh"

namespace hl {

constexpr int kStStages = 3;
constexpr int kStXFloats = 12288;       // x window per stage: 48 KB = 192 / 96 / 48 rows for F = 64 / 128 / 256
constexpr int kStMaxRows = 384;         // F = 32: 384 rows
constexpr int kStNzCap = 1536;          // colidx / vals entries staged per tile (larger tiles read them from global)
constexpr int kStConsumerWarps = 16;
constexpr int kStThreads = (kStConsumerWarps + 1) * 32;

struct StagedBatch {
  hl_spmm_problem p[HL_MAX_SPMM_PROBLEMS];
  int32_t tile_start[HL_MAX_SPMM_PROBLEMS + 1];
  int32_t n;
};

struct StageMeta {
  int32_t prob, row0, nrows, p0a, staged;
};

struct __align__(16) StagedSmem {
  float x[kStStages][kStXFloats];
  int32_t col[kStStages][kStNzCap + 8];
  float val[kStStages][kStNzCap + 8];
  int32_t rowptr[kStStages][kStMaxRows + 8];
  StageMeta meta[kStStages];
  unsigned long long full_bar[kStStages];
  unsigned long long empty_bar[kStStages];
};

__device__ __forceinline__ bool aligned_to_dev(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TileDesc {
  int32_t prob, row0, nrows;
};

__device__ __forceinline__ TileDesc tile_desc(const StagedBatch& b, int tile, int tile_rows) {
  int pb = 0;
  while (pb + 1 < b.n && tile >= b.tile_start[pb + 1]) ++pb;
  TileDesc d;
  d.prob = pb;
  d.row0 = (tile - b.tile_start[pb]) * tile_rows;
  d.nrows = min(tile_rows, b.p[pb].nrows - d.row0);
  return d;
}

struct TileArgs {
  const float* xs;          // staged x window (shared)
  const int32_t* cs;        // staged colidx, biased so that cs[q] is entry q of the operator (shared)
  const float* vs;          // staged vals, same bias (shared)
  const int32_t* rps;       // staged rowptr slice (shared)
  const int32_t* gcol; const float* gval; const float* xg;
  const float* p1; int64_t ld_p1;          // nullptr: own row comes from the staged window
  const float* p2; int64_t ld_p2;
  const float* p3; int64_t ld_p3;
  float* out; int64_t ld_out;
  bool has_p1;
  int32_t row0, nrows, width, half, lane_col;
  float c0, c1, c2, c3;
};

__device__ __forceinline__ void axpy4(float* acc, float v, const float4& x) {
  acc[0] = __fadd_rn(acc[0], __fmul_rn(v, x.x));
  acc[1] = __fadd_rn(acc[1], __fmul_rn(v, x.y));
  acc[2] = __fadd_rn(acc[2], __fmul_rn(v, x.z));
  acc[3] = __fadd_rn(acc[3], __fmul_rn(v, x.w));
}

template <int EPI, bool STAGED>
__device__ __forceinline__ void consume_tile(const TileArgs& A, int first_row, int row_stride) {
  const float* __restrict__ xs = A.xs;
  const int width = A.width, half = A.half, lane_col = A.lane_col, row0 = A.row0, nrows = A.nrows;
  for (int r = first_row; r < nrows; r += row_stride) {
    const int start = A.rps[r], end = A.rps[r + 1];
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = start; p < end; p += 4) {
      int c[4];
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = min(p + u, end - 1);
        if (STAGED) { c[u] = A.cs[q]; v[u] = A.vs[q]; }
        else { c[u] = __ldg(A.gcol + q); v[u] = __ldg(A.gval + q); }
      }
      int wr[4];
      unsigned worst = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        wr[u] = c[u] - row0;
        worst = max(worst, (unsigned)wr[u]);
      }
      float4 x0[4], x1[4];
      if (worst < (unsigned)nrows) {                          // fast path: all four neighbours inside the window
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* src = xs + wr[u] * width + lane_col;
          x0[u] = *reinterpret_cast<const float4*>(src);
          x1[u] = *reinterpret_cast<const float4*>(src + half);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if ((unsigned)wr[u] < (unsigned)nrows) {
            const float* src = xs + wr[u] * width + lane_col;
            x0[u] = *reinterpret_cast<const float4*>(src);
            x1[u] = *reinterpret_cast<const float4*>(src + half);
          } else {
            const float* src = A.xg + (int64_t)c[u] * width + lane_col;
            x0[u] = __ldg(reinterpret_cast<const float4*>(src));
            x1[u] = __ldg(reinterpret_cast<const float4*>(src + half));
          }
        }
      }
      const int left = end - p;
      axpy4(a0, v[0], x0[0]); axpy4(a1, v[0], x1[0]);
      if (left > 1) { axpy4(a0, v[1], x0[1]); axpy4(a1, v[1], x1[1]); }
      if (left > 2) { axpy4(a0, v[2], x0[2]); axpy4(a1, v[2], x1[2]); }
      if (left > 3) { axpy4(a0, v[3], x0[3]); axpy4(a1, v[3], x1[3]); }
    }
    // ---- recurrence epilogue (same arithmetic as poly_spmm_kernel) ----
    const int64_t grow = (int64_t)(row0 + r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* acc = h == 0 ? a0 : a1;
      const int col = lane_col + h * half;
      float o[4];
      float q1[4] = {0.f, 0.f, 0.f, 0.f}, q2[4] = {0.f, 0.f, 0.f, 0.f}, q3[4] = {0.f, 0.f, 0.f, 0.f};
      const bool need1 = EPI == HL_EPI_LAGUERRE_FIRST || EPI == HL_EPI_LAGUERRE_STEP || (EPI == HL_EPI_LINCOMB && A.has_p1);
      const bool need2 = EPI == HL_EPI_LAGUERRE_STEP || EPI == HL_EPI_CHEB_STEP || (EPI == HL_EPI_LINCOMB && A.p2);
      const bool need3 = EPI == HL_EPI_LINCOMB && A.p3;
      if (need1) {
        float4 t;
        if (A.p1 == nullptr) t = *reinterpret_cast<const float4*>(xs + r * width + col);
        else t = *reinterpret_cast<const float4*>(A.p1 + grow * A.ld_p1 + col);
        q1[0] = t.x; q1[1] = t.y; q1[2] = t.z; q1[3] = t.w;
      }
      if (need2) {
        const float4 t = *reinterpret_cast<const float4*>(A.p2 + grow * A.ld_p2 + col);
        q2[0] = t.x; q2[1] = t.y; q2[2] = t.z; q2[3] = t.w;
      }
      if (need3) {
        const float4 t = *reinterpret_cast<const float4*>(A.p3 + grow * A.ld_p3 + col);
        q3[0] = t.x; q3[1] = t.y; q3[2] = t.z; q3[3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (EPI == HL_EPI_CHEB_FIRST) o[i] = acc[i];
        else if (EPI == HL_EPI_LAGUERRE_FIRST) o[i] = __fsub_rn(q1[i], acc[i]);
        else if (EPI == HL_EPI_LAGUERRE_STEP) {
          float t = __fadd_rn(-acc[i], __fmul_rn(A.c1, q1[i]));
          t = __fsub_rn(t, __fmul_rn(A.c0, q2[i]));
          o[i] = __fdiv_rn(t, A.c2);
        } else if (EPI == HL_EPI_CHEB_STEP) o[i] = __fsub_rn(__fmul_rn(2.f, acc[i]), q2[i]);
        else {
          float t = A.c0 * acc[i];
          if (need1) t = fmaf(A.c1, q1[i], t);
          if (need2) t = fmaf(A.c2, q2[i], t);
          if (need3) t = fmaf(A.c3, q3[i], t);
          o[i] = t;
        }
      }
      *reinterpret_cast<float4*>(A.out + grow * A.ld_out + col) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// One work item = one tile of `tile_rows` consecutive rows of one operator, full width (rows contiguous).
template <int EPI>
__global__ void __launch_bounds__(kStThreads, 1)
poly_spmm_staged_kernel(const StagedBatch b, const int32_t width, const int32_t tile_rows, const int32_t total_tiles,
                        const float c0, const float c1, const float c2, const float c3) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StagedSmem& S = *reinterpret_cast<StagedSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStStages; ++s) {
      mbar_init(&S.full_bar[s], 33);                        // lane 0's arrive.expect_tx + 32 producer lanes
      mbar_init(&S.empty_bar[s], kStConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kStConsumerWarps) {
    // ===================================== producer warp =====================================
    // software-pipelined: the nnz range [p0, p1) of the NEXT tile is fetched while this tile's copies
    // are issued, so the only global latency the producer ever waits on is hidden behind a ring slot
    int w = blockIdx.x;
    if (w >= total_tiles) return;
    TileDesc cur = tile_desc(b, w, tile_rows);
    int p0 = __ldg(b.p[cur.prob].rowptr + cur.row0), p1 = __ldg(b.p[cur.prob].rowptr + cur.row0 + cur.nrows);
    for (int it = 0; w < total_tiles; ++it) {
      const int s = it % kStStages;
      const uint32_t round = (uint32_t)(it / kStStages);
      const int wn = w + gridDim.x;
      TileDesc nxt = cur;
      int np0 = 0, np1 = 0;
      if (wn < total_tiles) {
        nxt = tile_desc(b, wn, tile_rows);
        np0 = __ldg(b.p[nxt.prob].rowptr + nxt.row0);
        np1 = __ldg(b.p[nxt.prob].rowptr + nxt.row0 + nxt.nrows);
      }
      mbar_wait(&S.empty_bar[s], (round & 1u) ^ 1u);
      const hl_spmm_problem& P = b.p[cur.prob];
      const uint32_t x_bytes = (uint32_t)cur.nrows * (uint32_t)width * 4u;
      // 16-byte aligned index windows; bulk path only when the padded window stays inside the arrays
      const int p0a = p0 & ~3;
      const int n_al = (p1 - p0a + 3) & ~3;
      const bool cv_bulk = n_al <= kStNzCap && (int64_t)p0a + n_al <= (int64_t)P.nnz_hint &&
                           aligned_to_dev(P.colidx, 16) && aligned_to_dev(P.vals, 16);
      const int rp_n = (cur.nrows + 1 + 3) & ~3;
      const bool rp_bulk = cur.row0 + rp_n <= P.nrows + 1 && aligned_to_dev(P.rowptr, 16) && (cur.row0 & 3) == 0;
      const bool staged = cv_bulk || (p1 - p0a) <= kStNzCap;
      if (lane == 0) {
        uint32_t tx = x_bytes;
        if (cv_bulk) tx += 8u * (uint32_t)n_al;
        if (rp_bulk) tx += 4u * (uint32_t)rp_n;
        mbar_arrive_expect_tx(&S.full_bar[s], tx);
        bulk_g2s(&S.x[s][0], P.xg + (int64_t)cur.row0 * P.ld_xg, x_bytes, &S.full_bar[s]);
        if (cv_bulk) {
          bulk_g2s(&S.col[s][0], P.colidx + p0a, 4u * (uint32_t)n_al, &S.full_bar[s]);
          bulk_g2s(&S.val[s][0], P.vals + p0a, 4u * (uint32_t)n_al, &S.full_bar[s]);
        }
        if (rp_bulk) bulk_g2s(&S.rowptr[s][0], P.rowptr + cur.row0, 4u * (uint32_t)rp_n, &S.full_bar[s]);
        StageMeta m;
        m.prob = cur.prob; m.row0 = cur.row0; m.nrows = cur.nrows; m.p0a = p0a; m.staged = staged ? 1 : 0;
        S.meta[s] = m;
      }
      __syncwarp();
      if (!rp_bulk)
        for (int r = lane; r <= cur.nrows; r += 32) S.rowptr[s][r] = __ldg(P.rowptr + cur.row0 + r);
      if (!cv_bulk && staged)
        for (int q = p0a + lane; q < p1; q += 32) {
          S.col[s][q - p0a] = __ldg(P.colidx + q);
          S.val[s][q - p0a] = __ldg(P.vals + q);
        }
      mbar_arrive(&S.full_bar[s]);                          // every lane releases its own shared-memory writes
      cur = nxt; p0 = np0; p1 = np1; w = wn;
    }
    return;
  }

  // ======================================= consumer warps =======================================
  // LPR = width / 8 lanes per row, 2 float4 per lane (columns c and c + width/2)
  const int lpr = width >> 3;
  const int rows_per_warp = 32 / lpr;
  const int g = lane / lpr, gl = lane - g * lpr;
  const int half = width >> 1;
  const int lane_col = gl * 4;
  int it = 0;
  for (int w = blockIdx.x; w < total_tiles; w += gridDim.x, ++it) {
    const int s = it % kStStages;
    const uint32_t round = (uint32_t)(it / kStStages);
    mbar_wait(&S.full_bar[s], round & 1u);
    const StageMeta m = S.meta[s];
    const hl_spmm_problem& P = b.p[m.prob];
    const float* __restrict__ xs = &S.x[s][0];
    const int32_t* __restrict__ cs = &S.col[s][0];
    const float* __restrict__ vs = &S.val[s][0];
    const int32_t* __restrict__ rps = &S.rowptr[s][0];
    const bool own_in_smem = (P.p1 == P.xg) && (P.ld_p1 == P.ld_xg);

    // hoist everything the row loop needs out of the (dynamically indexed) parameter struct
    TileArgs A;
    A.xs = xs; A.cs = cs - m.p0a; A.vs = vs - m.p0a; A.rps = rps;
    A.gcol = P.colidx; A.gval = P.vals; A.xg = P.xg;
    A.p1 = own_in_smem ? nullptr : P.p1; A.ld_p1 = P.ld_p1;
    A.p2 = P.p2; A.ld_p2 = P.ld_p2; A.p3 = P.p3; A.ld_p3 = P.ld_p3;
    A.out = P.out; A.ld_out = P.ld_out;
    A.has_p1 = P.p1 != nullptr;
    A.row0 = m.row0; A.nrows = m.nrows; A.width = width; A.half = half; A.lane_col = lane_col;
    A.c0 = c0; A.c1 = c1; A.c2 = c2; A.c3 = c3;
    const int first = warp * rows_per_warp + g, stride = kStConsumerWarps * rows_per_warp;
    if (m.staged) consume_tile<EPI, true>(A, first, stride);
    else consume_tile<EPI, false>(A, first, stride);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty_bar[s]);
  }
}

// returns HL_OK if launched, 1 if this path does not apply (caller falls back to the per-row kernel)
int launch_poly_spmm_staged(const hl_spmm_problem* probs, int n, int32_t width, int epi, float c0, float c1, float c2,
                            float c3, cudaStream_t stream) {
  if (width != 32 && width != 64 && width != 128 && width != 256) return 1;
  int64_t rows = 0, nnz = 0;
  for (int i = 0; i < n; ++i) {
    const hl_spmm_problem& P = probs[i];
    const void* ptrs[5] = {P.xg, P.out, P.p1, P.p2, P.p3};
    const int64_t lds[5] = {P.ld_xg, P.ld_out, P.ld_p1, P.ld_p2, P.ld_p3};
    for (int k = 0; k < 5; ++k)
      if (ptrs[k] && (lds[k] % 4 != 0 || !aligned_to(ptrs[k], 16))) return 1;
    if (P.ld_xg != width) return 1;                           // window rows must be contiguous (one bulk copy)
    if (P.out == P.xg) return 1;
    rows += P.nrows;
    if (P.nnz_hint <= 0) return 1;                            // exact nnz unknown: keep the per-row kernel
    nnz += P.nnz_hint;
  }
  const int tile_rows = kStXFloats / width;
  if (rows < 4 * tile_rows) return 1;                        // tiny launches: the per-row kernel has less fixed cost
  if (nnz > 8 * rows) return 1;                              // long rows: window hit rate is low, indices do not fit
  StagedBatch b;
  b.n = n;
  int32_t tiles = 0;
  for (int i = 0; i < n; ++i) {
    b.p[i] = probs[i];
    b.tile_start[i] = tiles;
    tiles += (probs[i].nrows + tile_rows - 1) / tile_rows;
  }
  for (int i = n; i <= HL_MAX_SPMM_PROBLEMS; ++i) b.tile_start[i] = tiles;
  const int sm_count = device_sm_count();
  const int grid = tiles < sm_count ? tiles : sm_count;
  const size_t smem = sizeof(StagedSmem) + 128;
#define HL_ST_CASE(E)                                                                                         \
  case E: {                                                                                                   \
    static DeviceOnce configured;                                                                             \
    if (configured.need())                                                                                    \
      cudaFuncSetAttribute(poly_spmm_staged_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    hl::launch_pdl(poly_spmm_staged_kernel<E>, grid, kStThreads, smem, stream, b, width, tile_rows, tiles, c0, c1, c2, c3); \
  } break;
  switch (epi) {
    HL_ST_CASE(HL_EPI_LAGUERRE_FIRST)
    HL_ST_CASE(HL_EPI_LAGUERRE_STEP)
    HL_ST_CASE(HL_EPI_CHEB_FIRST)
    HL_ST_CASE(HL_EPI_CHEB_STEP)
    HL_ST_CASE(HL_EPI_LINCOMB)
    default: return 1;
  }
#undef HL_ST_CASE
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return record_cuda_error(e, "poly_spmm_staged_kernel");
  return HL_OK;
}

}  // namespace hl
