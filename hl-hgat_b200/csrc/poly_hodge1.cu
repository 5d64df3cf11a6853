// Factored application of the Hodge 1-Laplacian inside the polynomial recurrence:
//     L1 = diag(s) B1^T B1,   s[e] = 2 / lambda_max(graph of e)         (lib/Hodge_Dataset.py:456, no B2 term)
//     (L1 x)[e] = s[e] * (y[head_e] - y[tail_e]),   y[n] = sum_{f incident to n} sgn(n,f) x[f]
// i.e. 2 row gathers per node-incidence + 2 per edge instead of one per nonzero of L1 (deg(tail)+deg(head)-1:
// ~18 on CIFAR-superpixel graphs, ~55 on TSP kNN-25 graphs), where the CSR kernel is bound by the L1/L2 gather
// path rather than HBM (profiles/experiments/README.md).  The node intermediate y [N, F] is tiny next to the edge
// operands.  Opt-in (functional.enable_factored_hodge1): the caller asserts that the edge operator IS the
// Hodge 1-Laplacian of the batch's boundary matrix; results equal the CSR path up to fp32 summation order
// (~1e-7 relative), not bit for bit.  Same recurrence epilogues as hl_poly_spmm, no atomics, deterministic.
#include "common.cuh"

namespace hl {

// y[n,:] = sum over incident edges f (ascending id) of sgn(n,f) * x[f,:]; sgn = +1 at the head, -1 at the tail
template <int V>
__global__ void __launch_bounds__(256)
hodge1_node_kernel(const int32_t* __restrict__ inc_rowptr, const int32_t* __restrict__ inc_edge,
                   const int32_t* __restrict__ tail, const int32_t* __restrict__ head, int32_t n_nodes,
                   const float* __restrict__ x, int64_t ld_x, float* __restrict__ y, int32_t width, int32_t G) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int rows_per_block = 256 / G;
  const int gl = threadIdx.x & (G - 1);
  const int n = blockIdx.x * rows_per_block + (int)(threadIdx.x / G);
  if (n >= n_nodes) return;
  const int start = __ldg(inc_rowptr + n), end = __ldg(inc_rowptr + n + 1);
  for (int col = gl * V; col < width; col += G * V) {
    Pack<V> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
    for (int p = start; p < end; p += 4) {
      int f[4];
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        f[u] = __ldg(inc_edge + min(p + u, end - 1));
        const int t = __ldg(tail + f[u]), h = __ldg(head + f[u]);
        s[u] = (p + u < end && t != h) ? (h == n ? 1.f : -1.f) : 0.f;      // self-pairs (ghost padding): zero column of B1
      }
      Pack<V> xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = ld_pack<V>(x + (int64_t)f[u] * ld_x + col);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (s[u] != 0.f)
#pragma unroll
          for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], s[u] * xv[u].v[i]);
    }
    st_pack<V>(y + (int64_t)n * width + col, acc);
  }
}

struct Hodge1Epi {
  const float* p1; int64_t ld_p1;
  const float* p2; int64_t ld_p2;
  const float* p3; int64_t ld_p3;
  float* out; int64_t ld_out;
  float c0, c1, c2, c3;
};

// a = scale[e] * (y[head] - y[tail]); out = epilogue(a, own-row operands)
template <int V, int EPI>
__global__ void __launch_bounds__(256)
hodge1_edge_kernel(const int32_t* __restrict__ tail, const int32_t* __restrict__ head, int32_t n_edges,
                   const float* __restrict__ scale, const float* __restrict__ y, int32_t width, int32_t chunks,
                   const Hodge1Epi E) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int64_t total = (int64_t)n_edges * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx / chunks);
    const int col = (int)(idx - (int64_t)e * chunks) * V;
    const float sc = __ldg(scale + e);
    const Pack<V> yt = ld_pack<V>(y + (int64_t)__ldg(tail + e) * width + col);
    const Pack<V> yh = ld_pack<V>(y + (int64_t)__ldg(head + e) * width + col);
    const bool need1 = EPI == HL_EPI_LAGUERRE_FIRST || EPI == HL_EPI_LAGUERRE_STEP || (EPI == HL_EPI_LINCOMB && E.p1);
    const bool need2 = EPI == HL_EPI_LAGUERRE_STEP || EPI == HL_EPI_CHEB_STEP || (EPI == HL_EPI_LINCOMB && E.p2);
    const bool need3 = EPI == HL_EPI_LINCOMB && E.p3;
    Pack<V> q1, q2, q3, o;
#pragma unroll
    for (int i = 0; i < V; ++i) q1.v[i] = q2.v[i] = q3.v[i] = 0.f;
    if (need1) q1 = ld_pack_coherent<V>(E.p1 + (int64_t)e * E.ld_p1 + col);
    if (need2) q2 = ld_pack_coherent<V>(E.p2 + (int64_t)e * E.ld_p2 + col);
    if (need3) q3 = ld_pack_coherent<V>(E.p3 + (int64_t)e * E.ld_p3 + col);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = sc * (yh.v[i] - yt.v[i]);
      float r;
      if (EPI == HL_EPI_CHEB_FIRST) r = a;
      else if (EPI == HL_EPI_LAGUERRE_FIRST) r = __fsub_rn(q1.v[i], a);
      else if (EPI == HL_EPI_LAGUERRE_STEP) {
        float t = __fadd_rn(-a, __fmul_rn(E.c1, q1.v[i]));
        t = __fsub_rn(t, __fmul_rn(E.c0, q2.v[i]));
        r = __fdiv_rn(t, E.c2);
      } else if (EPI == HL_EPI_CHEB_STEP) r = __fsub_rn(__fmul_rn(2.f, a), q2.v[i]);
      else {
        r = E.c0 * a;
        if (need1) r = fmaf(E.c1, q1.v[i], r);
        if (need2) r = fmaf(E.c2, q2.v[i], r);
        if (need3) r = fmaf(E.c3, q3.v[i], r);
      }
      o.v[i] = r;
    }
    st_pack<V>(E.out + (int64_t)e * E.ld_out + col, o);
  }
}

static int hodge1_apply(const hl_hodge1_operator& op, const float* xg, int64_t ld_xg, float* node_tmp, int32_t width, int epi,
                        Hodge1Epi E, cudaStream_t st) {
  if (op.n_edges == 0) return HL_OK;
  int V = vec_for(xg, ld_xg, width, 4);
  V = min(V, vec_for(node_tmp, width, width, V));
  V = min(V, vec_for(E.out, E.ld_out, width, V));
  V = min(V, vec_for(E.p1, E.ld_p1, width, V));
  V = min(V, vec_for(E.p2, E.ld_p2, width, V));
  V = min(V, vec_for(E.p3, E.ld_p3, width, V));
  if (V == 2) V = 1;
  const int G = group_lanes(width, V);
  const int gridn = (op.n_nodes + 256 / G - 1) / (256 / G);
  if (op.n_nodes > 0) {
    if (V == 4) hl::launch_pdl(hodge1_node_kernel<4>, gridn, 256, 0, st, op.inc_rowptr, op.inc_edge, op.tail, op.head, op.n_nodes, xg, ld_xg, node_tmp, width, G);
    else hl::launch_pdl(hodge1_node_kernel<1>, gridn, 256, 0, st, op.inc_rowptr, op.inc_edge, op.tail, op.head, op.n_nodes, xg, ld_xg, node_tmp, width, G);
    HL_LAUNCH_CHECK("hodge1_node_kernel");
  }
  const int chunks = width / V;
  int64_t blocks = ((int64_t)op.n_edges * chunks + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
#define HL_H1_CASE(EP)                                                                                               \
  case EP:                                                                                                           \
    if (V == 4) hl::launch_pdl(hodge1_edge_kernel<4, EP>, (int)blocks, 256, 0, st, op.tail, op.head, op.n_edges, op.edge_scale, node_tmp, width, chunks, E); \
    else hl::launch_pdl(hodge1_edge_kernel<1, EP>, (int)blocks, 256, 0, st, op.tail, op.head, op.n_edges, op.edge_scale, node_tmp, width, chunks, E);        \
    break;
  switch (epi) {
    HL_H1_CASE(HL_EPI_LAGUERRE_FIRST)
    HL_H1_CASE(HL_EPI_LAGUERRE_STEP)
    HL_H1_CASE(HL_EPI_CHEB_FIRST)
    HL_H1_CASE(HL_EPI_CHEB_STEP)
    HL_H1_CASE(HL_EPI_LINCOMB)
    default: return HL_ERR_INVALID;
  }
#undef HL_H1_CASE
  HL_LAUNCH_CHECK("hodge1_edge_kernel");
  return HL_OK;
}

static void h1_recurrence(int family, int k, float* a, float* b, float* c) {   // T_{k+1} = a_k A T_k + b_k T_k + c_k T_{k-1}
  if (family == HL_LAGUERRE) {
    if (k == 0) { *a = -1.f; *b = 1.f; *c = 0.f; }
    else { *a = -1.f / (k + 1); *b = (2.f * k + 1.f) / (k + 1); *c = -(float)k / (k + 1); }
  } else {
    if (k == 0) { *a = 1.f; *b = 0.f; *c = 0.f; }
    else { *a = 2.f; *b = 0.f; *c = -1.f; }
  }
}

static bool h1_valid(const hl_hodge1_operator* op) {
  return op && op->n_nodes >= 0 && op->n_edges >= 0 &&
         (op->n_edges == 0 || (op->inc_rowptr && op->inc_edge && op->tail && op->head && op->edge_scale));
}

}  // namespace hl

extern "C" int hl_poly_basis_hodge1_fwd(int family, int K, const hl_hodge1_operator* op, const float* x, int64_t ld_x,
                                        float* t, int64_t ld_t, int64_t t_stride, float* node_tmp, int32_t width,
                                        hl_stream_t stream) {
  using namespace hl;
  if ((family != HL_LAGUERRE && family != HL_CHEB) || K < 1 || width < 1 || !h1_valid(op)) return HL_ERR_INVALID;
  if (K > 1 && (!x || !t || !node_tmp)) return HL_ERR_INVALID;
  auto T = [&](int j) -> const float* { return j == 0 ? x : t + (int64_t)(j - 1) * t_stride; };
  auto LD = [&](int j) -> int64_t { return j == 0 ? ld_x : ld_t; };
  for (int k = 0; k + 1 < K; ++k) {
    Hodge1Epi E;
    E.p1 = T(k); E.ld_p1 = LD(k);
    E.p2 = k > 0 ? T(k - 1) : nullptr; E.ld_p2 = k > 0 ? LD(k - 1) : 0;
    E.p3 = nullptr; E.ld_p3 = 0;
    E.out = t + (int64_t)k * t_stride; E.ld_out = ld_t;
    E.c0 = (float)k; E.c1 = 2.f * k + 1.f; E.c2 = k + 1.f; E.c3 = 0.f;
    int epi;
    if (family == HL_LAGUERRE) epi = (k == 0) ? HL_EPI_LAGUERRE_FIRST : HL_EPI_LAGUERRE_STEP;
    else epi = (k == 0) ? HL_EPI_CHEB_FIRST : HL_EPI_CHEB_STEP;
    const int rc = hodge1_apply(*op, T(k), LD(k), node_tmp, width, epi, E, as_stream(stream));
    if (rc != HL_OK) return rc;
  }
  return HL_OK;
}

extern "C" int hl_poly_basis_hodge1_bwd(int family, int K, const hl_hodge1_operator* op, float* g0, int64_t ld_g0,
                                        float* t, int64_t ld_t, int64_t t_stride, float* node_tmp, int32_t width,
                                        hl_stream_t stream) {
  using namespace hl;
  if ((family != HL_LAGUERRE && family != HL_CHEB) || K < 1 || width < 1 || !h1_valid(op) || !g0) return HL_ERR_INVALID;
  if (K > 1 && (!t || !node_tmp)) return HL_ERR_INVALID;
  auto Gp = [&](int j) -> float* { return j == 0 ? g0 : t + (int64_t)(j - 1) * t_stride; };
  auto LD = [&](int j) -> int64_t { return j == 0 ? ld_g0 : ld_t; };
  // S_{K-1} = G_{K-1};  S_k = G_k + a_k L1 S_{k+1} + b_k S_{k+1} + c_{k+1} S_{k+2}, in place (L1 is symmetric)
  for (int k = K - 2; k >= 0; --k) {
    float ak, bk, ck, a1, b1, ck1 = 0.f;
    (void)ck; (void)a1; (void)b1;
    h1_recurrence(family, k, &ak, &bk, &ck);
    if (k + 2 <= K - 1) h1_recurrence(family, k + 1, &a1, &b1, &ck1);
    Hodge1Epi E;
    E.p1 = (bk != 0.f) ? Gp(k + 1) : nullptr; E.ld_p1 = LD(k + 1);
    E.p2 = (k + 2 <= K - 1 && ck1 != 0.f) ? Gp(k + 2) : nullptr; E.ld_p2 = (k + 2 <= K - 1) ? LD(k + 2) : 0;
    E.p3 = Gp(k); E.ld_p3 = LD(k);
    E.out = Gp(k); E.ld_out = LD(k);
    E.c0 = ak; E.c1 = bk; E.c2 = ck1; E.c3 = 1.f;
    const int rc = hodge1_apply(*op, Gp(k + 1), LD(k + 1), node_tmp, width, HL_EPI_LINCOMB, E, as_stream(stream));
    if (rc != HL_OK) return rc;
  }
  return HL_OK;
}
