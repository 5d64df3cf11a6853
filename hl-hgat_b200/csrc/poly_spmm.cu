// Polynomial CSR SpMM with fused recurrence epilogue (hl_poly_spmm, hl_poly_basis_fwd/bwd).
//
// HBM-bound gather kernel.  A group of G lanes (power of two, G*V*CH >= tile width) owns one output
// row; the group streams the row's (colidx, vals) coalesced, broadcasts them by shuffle and gathers
// V-wide (128-bit for V = 4) slices of the source rows.  Each output element is accumulated serially in
// ascending CSR order with the product rounded before the add, which reproduces the CPU reference's
// index_add_ summation bit for bit ("Deterministic ordering").  The recurrence
// (Laguerre / Chebyshev step or a general linear combination for the adjoint) is applied in registers
// before the single store, so T_{k+1} never makes an extra round trip through HBM.  Up to four
// operators (node-side L0 and edge-side L1 of a block) share one launch.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace hl {

struct SpmmBatch {
  hl_spmm_problem p[HL_MAX_SPMM_PROBLEMS];
  int32_t block_start[HL_MAX_SPMM_PROBLEMS + 1];
  int32_t n;
};

constexpr int kThreads = 256;

// One group of G lanes per output row; each lane owns CH vectors of V floats (columns
// col0 + ch*G*V).  The row's (colidx, vals) are read with group-uniform (broadcast) loads, U entries
// at a time: all U*CH gathers are issued before the ordered multiply/add chain consumes them, which
// is where the memory-level parallelism comes from on short rows (ZINC: 3-4 nnz per row).
template <int V, int CH, int U, int EPI>
__global__ void __launch_bounds__(kThreads)
poly_spmm_kernel(const SpmmBatch b, const int32_t width, const int32_t G,
                 const float c0, const float c1, const float c2, const float c3) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  int pb = 0;
  while (pb + 1 < b.n && (int32_t)blockIdx.x >= b.block_start[pb + 1]) ++pb;
  const hl_spmm_problem& P = b.p[pb];

  const int rows_per_block = kThreads / G;
  const int gl = threadIdx.x & (G - 1);                       // lane within the row group
  const int row = ((int32_t)blockIdx.x - b.block_start[pb]) * rows_per_block + (int)(threadIdx.x / G);
  if (row >= P.nrows) return;

  const int col0 = blockIdx.y * (G * V * CH) + gl * V;        // first column of chunk 0
  bool act[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) act[ch] = (col0 + ch * G * V) < width;

  Pack<V> acc[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[ch].v[i] = 0.f;

  const int start = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  const float* __restrict__ xg = P.xg + col0;
  const int64_t ldx = P.ld_xg;
  const int32_t* __restrict__ colidx = P.colidx;
  const float* __restrict__ vals = P.vals;

  for (int p = start; p < end; p += U) {
    int c[U];
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = min(p + u, end - 1);                      // clamped: loads stay unconditional
      c[u] = __ldg(colidx + q);
      v[u] = __ldg(vals + q);
    }
    Pack<V> x[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        if (act[ch]) x[u][ch] = ld_pack<V>(xg + (int64_t)c[u] * ldx + ch * G * V);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u < end) {
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (act[ch])
#pragma unroll
            for (int i = 0; i < V; ++i)
              acc[ch].v[i] = __fadd_rn(acc[ch].v[i], __fmul_rn(v[u], x[u][ch].v[i]));
      }
  }

  // ---- fused recurrence epilogue -------------------------------------------------------------
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    if (!act[ch]) continue;
    const int col = col0 + ch * G * V;
    Pack<V> o;
    if constexpr (EPI == HL_EPI_CHEB_FIRST) {
      o = acc[ch];
    } else if constexpr (EPI == HL_EPI_LAGUERRE_FIRST) {
      Pack<V> a1 = ld_pack_coherent<V>(P.p1 + (int64_t)row * P.ld_p1 + col);
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = __fsub_rn(a1.v[i], acc[ch].v[i]);
    } else if constexpr (EPI == HL_EPI_LAGUERRE_STEP) {
      // (-a + (2k+1) T_k - k T_{k-1}) / (k+1), in the reference's evaluation order
      Pack<V> a1 = ld_pack_coherent<V>(P.p1 + (int64_t)row * P.ld_p1 + col);
      Pack<V> a2 = ld_pack_coherent<V>(P.p2 + (int64_t)row * P.ld_p2 + col);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float t = __fadd_rn(-acc[ch].v[i], __fmul_rn(c1, a1.v[i]));
        t = __fsub_rn(t, __fmul_rn(c0, a2.v[i]));
        o.v[i] = __fdiv_rn(t, c2);
      }
    } else if constexpr (EPI == HL_EPI_CHEB_STEP) {
      Pack<V> a2 = ld_pack_coherent<V>(P.p2 + (int64_t)row * P.ld_p2 + col);
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = __fsub_rn(__fmul_rn(2.f, acc[ch].v[i]), a2.v[i]);
    } else {  // HL_EPI_LINCOMB
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = c0 * acc[ch].v[i];
      if (P.p1) {
        Pack<V> a1 = ld_pack_coherent<V>(P.p1 + (int64_t)row * P.ld_p1 + col);
#pragma unroll
        for (int i = 0; i < V; ++i) o.v[i] = fmaf(c1, a1.v[i], o.v[i]);
      }
      if (P.p2) {
        Pack<V> a2 = ld_pack_coherent<V>(P.p2 + (int64_t)row * P.ld_p2 + col);
#pragma unroll
        for (int i = 0; i < V; ++i) o.v[i] = fmaf(c2, a2.v[i], o.v[i]);
      }
      if (P.p3) {
        Pack<V> a3 = ld_pack_coherent<V>(P.p3 + (int64_t)row * P.ld_p3 + col);
#pragma unroll
        for (int i = 0; i < V; ++i) o.v[i] = fmaf(c3, a3.v[i], o.v[i]);
      }
    }
    st_pack<V>(P.out + (int64_t)row * P.ld_out + col, o);
  }
}

template <int V, int CH, int U>
static void launch_epi(dim3 grid, cudaStream_t stream, const SpmmBatch& b, int32_t width, int G, int epi,
                       float c0, float c1, float c2, float c3) {
  switch (epi) {
#define HL_EPI_CASE(E) \
  case E: hl::launch_pdl(poly_spmm_kernel<V, CH, U, E>, grid, kThreads, 0, stream, b, width, G, c0, c1, c2, c3); break;
    HL_EPI_CASE(HL_EPI_LAGUERRE_FIRST)
    HL_EPI_CASE(HL_EPI_LAGUERRE_STEP)
    HL_EPI_CASE(HL_EPI_CHEB_FIRST)
    HL_EPI_CASE(HL_EPI_CHEB_STEP)
    default: hl::launch_pdl(poly_spmm_kernel<V, CH, U, HL_EPI_LINCOMB>, grid, kThreads, 0, stream, b, width, G, c0, c1, c2, c3);
#undef HL_EPI_CASE
  }
}

int launch_poly_spmm_staged(const hl_spmm_problem* probs, int n, int32_t width, int epi, float c0, float c1, float c2,
                            float c3, cudaStream_t stream);   // poly_spmm_staged.cu

// HL_SPMM_MODE = auto (default) | rows (per-row kernel only) | staged (staged kernel whenever it applies)
static int g_spmm_mode = -1;
static int spmm_mode() {
  if (g_spmm_mode < 0) {
    const char* e = getenv("HL_SPMM_MODE");
    g_spmm_mode = (e && !strcmp(e, "rows")) ? 1 : 0;
  }
  return g_spmm_mode;
}

static int launch_poly_spmm(const hl_spmm_problem* probs, int n, int32_t width, int epi, const float* c,
                            cudaStream_t stream) {
  if (!probs || n < 1 || n > HL_MAX_SPMM_PROBLEMS || width < 1) return HL_ERR_INVALID;
  if (epi < HL_EPI_LAGUERRE_FIRST || epi > HL_EPI_LINCOMB) return HL_ERR_INVALID;
  int V = 4;
  for (int i = 0; i < n; ++i) {
    const hl_spmm_problem& P = probs[i];
    if (P.nrows < 0 || !P.out || !P.xg || (P.nrows > 0 && (!P.rowptr || !P.colidx || !P.vals))) return HL_ERR_INVALID;
    if ((epi == HL_EPI_LAGUERRE_FIRST || epi == HL_EPI_LAGUERRE_STEP) && !P.p1) return HL_ERR_INVALID;
    if ((epi == HL_EPI_LAGUERRE_STEP || epi == HL_EPI_CHEB_STEP) && !P.p2) return HL_ERR_INVALID;
    V = min(V, vec_for(P.xg, P.ld_xg, width, V));
    V = min(V, vec_for(P.out, P.ld_out, width, V));
    V = min(V, vec_for(P.p1, P.ld_p1, width, V));
    V = min(V, vec_for(P.p2, P.ld_p2, width, V));
    V = min(V, vec_for(P.p3, P.ld_p3, width, V));
  }
  if (V == 4 && spmm_mode() == 0) {
    float s0 = c ? c[0] : 0.f, s1 = c ? c[1] : 0.f, s2 = c ? c[2] : 0.f, s3 = c ? c[3] : 0.f;
    if (epi == HL_EPI_LAGUERRE_STEP) { const float k = s0; s0 = k; s1 = 2.f * k + 1.f; s2 = k + 1.f; }
    const int rc = launch_poly_spmm_staged(probs, n, width, epi, s0, s1, s2, s3, stream);
    if (rc != 1) return rc;                                  // 1 = not applicable, fall through
  }
  // two vectors per lane whenever the row is wide enough to keep >= 8 lanes busy: halves the
  // per-row index/address instruction overhead and every lane still reads full 128-byte lines
  const int chunks = (width + V - 1) / V;
  const int CH = chunks >= 16 ? 2 : 1;                       // (at 8 chunks two vectors per lane would touch half lines)
  int G = 1;
  while (G * CH < chunks && G < 32) G <<= 1;
  const int tile_w = G * V * CH;
  const int rows_per_block = kThreads / G;

  SpmmBatch b;
  b.n = n;
  int32_t blocks = 0;
  for (int i = 0; i < n; ++i) {
    b.p[i] = probs[i];
    b.block_start[i] = blocks;
    blocks += (probs[i].nrows + rows_per_block - 1) / rows_per_block;
  }
  for (int i = n; i <= HL_MAX_SPMM_PROBLEMS; ++i) b.block_start[i] = blocks;
  if (blocks == 0) return HL_OK;
  dim3 grid(blocks, (width + tile_w - 1) / tile_w);
  float c0 = c ? c[0] : 0.f, c1 = c ? c[1] : 0.f, c2 = c ? c[2] : 0.f, c3 = c ? c[3] : 0.f;
  if (epi == HL_EPI_LAGUERRE_STEP) {                       // c[0] = k -> (k, 2k+1, k+1)
    const float k = c0;
    c0 = k; c1 = 2.f * k + 1.f; c2 = k + 1.f;
  }
  if (V == 4) { if (CH == 2) launch_epi<4, 2, 2>(grid, stream, b, width, G, epi, c0, c1, c2, c3);
                else launch_epi<4, 1, 4>(grid, stream, b, width, G, epi, c0, c1, c2, c3); }
  else if (V == 2) { if (CH == 2) launch_epi<2, 2, 4>(grid, stream, b, width, G, epi, c0, c1, c2, c3);
                     else launch_epi<2, 1, 4>(grid, stream, b, width, G, epi, c0, c1, c2, c3); }
  else { if (CH == 2) launch_epi<1, 2, 4>(grid, stream, b, width, G, epi, c0, c1, c2, c3);
         else launch_epi<1, 1, 4>(grid, stream, b, width, G, epi, c0, c1, c2, c3); }
  HL_LAUNCH_CHECK("poly_spmm_kernel");
  return HL_OK;
}

// coefficients of T_{k+1} = a_k A T_k + b_k T_k + c_k T_{k-1}
static void recurrence(int family, int k, float* a, float* b, float* c) {
  if (family == HL_LAGUERRE) {
    if (k == 0) { *a = -1.f; *b = 1.f; *c = 0.f; }
    else { *a = -1.f / (k + 1); *b = (2.f * k + 1.f) / (k + 1); *c = -(float)k / (k + 1); }
  } else {
    if (k == 0) { *a = 1.f; *b = 0.f; *c = 0.f; }
    else { *a = 2.f; *b = 0.f; *c = -1.f; }
  }
}

}  // namespace hl

extern "C" void hl_set_spmm_mode(int mode) { hl::g_spmm_mode = mode ? 1 : 0; }

extern "C" int hl_poly_spmm(const hl_spmm_problem* problems, int nproblems, int32_t width, int epilogue,
                            const float* c, hl_stream_t stream) {
  return hl::launch_poly_spmm(problems, nproblems, width, epilogue, c, hl::as_stream(stream));
}

extern "C" int hl_poly_basis_fwd(int family, int K, const hl_conv_side* sides, int nsides, int32_t width,
                                 hl_stream_t stream) {
  if ((family != HL_LAGUERRE && family != HL_CHEB) || K < 1 || !sides || nsides < 1 ||
      nsides > HL_MAX_SPMM_PROBLEMS)
    return HL_ERR_INVALID;
  hl_spmm_problem pr[HL_MAX_SPMM_PROBLEMS];
  for (int k = 0; k + 1 < K; ++k) {                         // produce T_{k+1}
    for (int s = 0; s < nsides; ++s) {
      const hl_conv_side& S = sides[s];
      if (!S.t || !S.x) return HL_ERR_INVALID;
      auto T = [&](int j) -> const float* { return j == 0 ? S.x : S.t + (int64_t)(j - 1) * S.t_stride; };
      auto LD = [&](int j) -> int64_t { return j == 0 ? S.ld_x : S.ld_t; };
      hl_spmm_problem& P = pr[s];
      P.rowptr = S.rowptr; P.colidx = S.colidx; P.vals = S.vals; P.nrows = S.nrows; P.nnz_hint = S.nnz_hint;
      P.xg = T(k); P.ld_xg = LD(k);
      P.p1 = T(k); P.ld_p1 = LD(k);
      P.p2 = k > 0 ? T(k - 1) : nullptr; P.ld_p2 = k > 0 ? LD(k - 1) : 0;
      P.p3 = nullptr; P.ld_p3 = 0;
      P.out = S.t + (int64_t)k * S.t_stride; P.ld_out = S.ld_t;
    }
    int epi;
    float c[4] = {(float)k, 0.f, 0.f, 0.f};
    if (family == HL_LAGUERRE) epi = (k == 0) ? HL_EPI_LAGUERRE_FIRST : HL_EPI_LAGUERRE_STEP;
    else epi = (k == 0) ? HL_EPI_CHEB_FIRST : HL_EPI_CHEB_STEP;
    int rc = hl::launch_poly_spmm(pr, nsides, width, epi, c, hl::as_stream(stream));
    if (rc != HL_OK) return rc;
  }
  return HL_OK;
}

extern "C" int hl_poly_basis_bwd(int family, int K, const hl_conv_side* sides, int nsides, int32_t width,
                                 hl_stream_t stream) {
  if ((family != HL_LAGUERRE && family != HL_CHEB) || K < 1 || !sides || nsides < 1 ||
      nsides > HL_MAX_SPMM_PROBLEMS)
    return HL_ERR_INVALID;
  hl_spmm_problem pr[HL_MAX_SPMM_PROBLEMS];
  // S_{K-1} = G_{K-1};  S_k = G_k + a_k A^T S_{k+1} + b_k S_{k+1} + c_{k+1} S_{k+2}, in place.
  for (int k = K - 2; k >= 0; --k) {
    float ak, bk, ck, a1, b1, ck1 = 0.f;
    hl::recurrence(family, k, &ak, &bk, &ck);
    if (k + 2 <= K - 1) hl::recurrence(family, k + 1, &a1, &b1, &ck1);
    for (int s = 0; s < nsides; ++s) {
      const hl_conv_side& S = sides[s];
      if (!S.g0 || (K > 1 && !S.t)) return HL_ERR_INVALID;
      auto Gp = [&](int j) -> float* { return j == 0 ? S.g0 : S.t + (int64_t)(j - 1) * S.t_stride; };
      auto LD = [&](int j) -> int64_t { return j == 0 ? S.ld_g0 : S.ld_t; };
      hl_spmm_problem& P = pr[s];
      P.rowptr = S.rowptr; P.colidx = S.colidx; P.vals = S.vals; P.nrows = S.nrows; P.nnz_hint = S.nnz_hint;
      P.xg = Gp(k + 1); P.ld_xg = LD(k + 1);
      P.p1 = (bk != 0.f) ? Gp(k + 1) : nullptr; P.ld_p1 = LD(k + 1);
      P.p2 = (k + 2 <= K - 1 && ck1 != 0.f) ? Gp(k + 2) : nullptr; P.ld_p2 = (k + 2 <= K - 1) ? LD(k + 2) : 0;
      P.p3 = Gp(k); P.ld_p3 = LD(k);
      P.out = Gp(k); P.ld_out = LD(k);
    }
    float c[4] = {ak, bk, ck1, 1.f};
    int rc = hl::launch_poly_spmm(pr, nsides, width, HL_EPI_LINCOMB, c, hl::as_stream(stream));
    if (rc != HL_OK) return rc;
  }
  return HL_OK;
}
