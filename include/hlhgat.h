/*
 * hlhgat.h -- C ABI of libhlhgat.so: the B200 (sm_100a) kernels behind the HL-HGAT hot path.
 *
 * The reference (deepika090/HL-HGAT) has no FFI of its own: every device kernel on this path is
 * reached through a third-party Python call (PyG MessagePassing.propagate, torch.sparse.mm,
 * torch_scatter.scatter_mean, ...).  Each entry point below names the reference call site(s) it
 * replaces (paths relative to the reference root).  Conventions for every function:
 *   - all data pointers are DEVICE pointers unless the parameter is documented as host;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - no allocation inside: scratch is passed in, sized by the matching *_workspace() query;
 *   - returns HL_OK (0) or a negative hl_status; never throws;
 *   - matrices are row-major fp32 with an explicit leading dimension `ld` (elements), so callers
 *     can read/write column slices of wider buffers (the dense-connection concat buffers);
 *   - indices inside the library are int32 (the reference uses int64 COO; conversion happens
 *     once per mini-batch in hl_csr_from_coo / hl_build_*).
 * Thread-compatible, not thread-safe.
 */
#ifndef HLHGAT_H_
#define HLHGAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HL_ABI_VERSION 1

typedef enum {
  HL_OK = 0,
  HL_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unsupported mode) */
  HL_ERR_WORKSPACE = -2,   /* workspace too small */
  HL_ERR_CUDA = -3,        /* a CUDA runtime call / launch failed: see hl_last_cuda_error() */
  HL_ERR_ALIGN = -4        /* pointer / leading dimension not aligned as documented */
} hl_status;

typedef void* hl_stream_t; /* cudaStream_t */

int hl_version(void);
const char* hl_status_string(int status);
const char* hl_last_cuda_error(void);
int hl_device_sm_count(void);
/* number of kernels this library has launched in this process (host-side counter) */
unsigned long long hl_launch_count(void);

/* --------------------------------------------------------------------------------------------
 * CSR construction from a COO edge list (once per mini-batch and operator).
 * Replaces: the implicit per-call work of PyG `propagate` (lib/Hodge_Cheb_Conv.py:494,502:
 * index_select + scatter over an int64 COO), the `par.abs()` re-coalesce of
 * lib/Hodge_Cheb_Conv.py:294-295, and the cluster bucketing inside torch_scatter.scatter_mean
 * (lib/Hodge_ST_Model.py:147,150).
 *
 * Groups entries by `row[i]` (rows outside [0,nrows) are dropped) and orders each group by
 *   tie == HL_TIE_POSITION : the entry's position i in the COO (the CPU reference's summation
 *                            order for index_add_; SURVEY.md section 7 "Deterministic ordering")
 *   tie == HL_TIE_COLUMN   : col[i] ascending (the order of a coalesced sparse tensor).
 * Outputs: rowptr[nrows+1], colidx[nnz_kept], vals[nnz_kept] (if val != NULL),
 *          perm[nnz_kept] = original position (if perm != NULL).  rowptr[nrows] = nnz_kept.
 * `row_f32`: alternative float32 row ids (the reference's float cluster ids, +inf = dropped,
 * lib/Hodge_ST_Model.py:142-144); exactly one of row / row_f32 must be non-NULL.
 * -------------------------------------------------------------------------------------------- */
enum { HL_TIE_POSITION = 0, HL_TIE_COLUMN = 1 };
size_t hl_csr_from_coo_workspace(int64_t nnz, int64_t nrows);
int hl_csr_from_coo(const int64_t* row, const float* row_f32, const int64_t* col /* NULL = position */,
                    const float* val, int64_t nnz, int64_t nrows, int tie,
                    int32_t* rowptr, int32_t* colidx, float* vals, int32_t* perm,
                    void* workspace, size_t workspace_bytes, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Polynomial SpMM with fused recurrence epilogue: one launch per polynomial order over one or
 * several operators (node side L0 and edge side L1 of a NEConv block share a launch).
 * Replaces: `self.propagate(edge_index, x=..., norm=...)` + the recurrence arithmetic,
 * lib/Hodge_Cheb_Conv.py:494 (T1 = x - A x), :502+:507 (Laguerre step), :412-416 (Chebyshev
 * T1), :430-432 (Chebyshev step); HL-HGAT-DEMO/lib/Hodge_Cheb_Conv.py:577-578 (CSR SpMM).
 *
 *   a[i,:]   = sum_{p in row i, ascending} vals[p] * xg[colidx[p], :]     (product rounded, then added:
 *                                                                          bit-equal to the CPU reference)
 *   out[i,:] = HL_EPI_LAGUERRE_FIRST : p1 - a
 *              HL_EPI_LAGUERRE_STEP  : (-a + (2k+1) p1 - k p2) / (k+1)     k = (int)c[0]
 *              HL_EPI_CHEB_FIRST     : a
 *              HL_EPI_CHEB_STEP      : 2 a - p2
 *              HL_EPI_LINCOMB        : c[0] a + c[1] p1 + c[2] p2 + c[3] p3  (NULL operands skipped)
 * `width` columns; every ld and pointer must allow the widest vector that divides `width`
 * (16 B for width % 4 == 0, 8 B for width % 2 == 0).  out may alias p1/p2/p3 (own-row only) but
 * not xg.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* vals;
  int32_t nrows;
  int32_t nnz_hint;                /* total nnz if known on the host (0 = unknown): selects the staged kernel */
  const float* xg;  int64_t ld_xg;
  const float* p1;  int64_t ld_p1;
  const float* p2;  int64_t ld_p2;
  const float* p3;  int64_t ld_p3;
  float* out;       int64_t ld_out;
} hl_spmm_problem;

enum {
  HL_EPI_LAGUERRE_FIRST = 0,
  HL_EPI_LAGUERRE_STEP = 1,
  HL_EPI_CHEB_FIRST = 2,
  HL_EPI_CHEB_STEP = 3,
  HL_EPI_LINCOMB = 4
};
#define HL_MAX_SPMM_PROBLEMS 4
/* kernel selection: 0 = auto (row-window staged kernel whenever it applies), 1 = per-row kernel only.
 * Also settable with the environment variable HL_SPMM_MODE=rows.  Both kernels give identical bits. */
void hl_set_spmm_mode(int mode);
int hl_poly_spmm(const hl_spmm_problem* problems /* host */, int nproblems, int32_t width, int epilogue,
                 const float* c /* host, 4 floats */, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Whole polynomial basis of one convolution (K-1 launches of hl_poly_spmm):
 * forward: block k-1 of t (at t + (k-1)*t_stride, leading dimension ld_t) = T_k(x), k = 1..K-1
 *          (T_0 = x is not copied).  t_stride = width gives the column-concatenated layout
 *          [R,(K-1)*width]; t_stride = nrows*ld_t the stacked layout [K-1,R,width].
 * backward: given g0 = dL/dT_0-direct [R,width] and block k-1 of t = dL/dT_k-direct, runs the
 *           adjoint recurrence in place on the TRANSPOSED operator; on return g0 holds dL/dx.
 * Replaces: HodgeLaguerreConv.forward / HodgeChebConv.forward minus the dense Linear layers
 * (lib/Hodge_Cheb_Conv.py:480-515, :394-439) and their autograd.
 * family: HL_LAGUERRE / HL_CHEB.  All sides share `width` (else call once per side).
 * -------------------------------------------------------------------------------------------- */
enum { HL_LAGUERRE = 0, HL_CHEB = 1 };
typedef struct {
  const int32_t* rowptr;   /* forward: CSR by target row; backward: CSR of the transpose */
  const int32_t* colidx;
  const float* vals;
  int32_t nrows;
  int32_t nnz_hint;        /* total nnz if known on the host, else 0 */
  const float* x;  int64_t ld_x;   /* forward: input x.   backward: unused (NULL) */
  float* t;        int64_t ld_t;   /* forward: basis out (K-1 blocks). backward: gt (in/out) */
  int64_t t_stride;                /* elements between consecutive blocks of t */
  float* g0;       int64_t ld_g0;  /* backward only: in dL/dT_0-direct, out dL/dx */
} hl_conv_side;
int hl_poly_basis_fwd(int family, int K, const hl_conv_side* sides /* host */, int nsides, int32_t width,
                      hl_stream_t stream);
int hl_poly_basis_bwd(int family, int K, const hl_conv_side* sides /* host */, int nsides, int32_t width,
                      hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Polynomial basis of the Hodge 1-Laplacian applied in FACTORED form (opt-in):
 *   L1 = diag(edge_scale) B1^T B1,  edge_scale[e] = 2 / lambda_max(graph of e)   (lib/Hodge_Dataset.py:456: no B2 term)
 *   (L1 x)[e] = edge_scale[e] * (y[head_e] - y[tail_e]),  y[n] = sum_{f incident to n} sgn(n,f) x[f]
 * 4 row gathers per edge instead of one per nonzero of L1 (deg(tail)+deg(head)-1: ~18 CIFAR-superpixel, ~55 TSP).
 * Same outputs as hl_poly_basis_fwd / _bwd on the CSR of that operator up to fp32 summation order (NOT bit for
 * bit); the caller asserts that the edge operator of the batch is this Laplacian.  node_tmp: [n_nodes, width]
 * scratch.  Replaces the same reference lines as hl_poly_basis_fwd/bwd for edge_index_s / edge_weight_s.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
  const int32_t* inc_rowptr;       /* node -> incident-edge CSR (ascending edge id) */
  const int32_t* inc_edge;
  const int32_t* tail;             /* [n_edges] */
  const int32_t* head;
  const float* edge_scale;         /* [n_edges] */
  int32_t n_nodes, n_edges;
} hl_hodge1_operator;
int hl_poly_basis_hodge1_fwd(int family, int K, const hl_hodge1_operator* op /* host */, const float* x, int64_t ld_x,
                             float* t, int64_t ld_t, int64_t t_stride, float* node_tmp, int32_t width, hl_stream_t stream);
int hl_poly_basis_hodge1_bwd(int family, int K, const hl_hodge1_operator* op /* host */, float* g0, int64_t ld_g0,
                             float* t, int64_t ld_t, int64_t t_stride, float* node_tmp, int32_t width, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Segmented (CSR-bucketed) row reduction with deterministic ascending order.
 *   dst[r,:] = post( sum_{p in [rowptr[r], rowptr[r+1])} pre(src[m_p, :]) ),  m_p = colidx ? colidx[p] : p
 *   pre  : src_scale ? src_scale[m] * v : v                  (attention gate, rounded before the sum)
 *   post : HL_POST_NONE     s
 *          HL_POST_CONST    s * cscale
 *          HL_POST_RCP_ROW  (1.0f / row_scale[r]) * s         (the 1/D of :294)
 *          HL_POST_MEAN     s / max(count, 1)                 (scatter_mean / global_mean_pool)
 * Replaces: torch.sparse.mm(par.abs(), x_s) * (1/D)  lib/Hodge_Cheb_Conv.py:294 (=:100);
 *           its adjoint |B1| g / 2 (backward of :295); scatter_mean lib/Hodge_ST_Model.py:147,150,
 *           lib/Hodge_Cheb_Conv.py:50,53; global_mean_pool lib/Hodge_ST_Model.py:636.
 * -------------------------------------------------------------------------------------------- */
enum { HL_POST_NONE = 0, HL_POST_CONST = 1, HL_POST_RCP_ROW = 2, HL_POST_MEAN = 3 };
int hl_segment_reduce(const int32_t* rowptr, const int32_t* colidx, int32_t nrows,
                      const float* src, int64_t ld_src, const float* src_scale,
                      float* dst, int64_t ld_dst, int32_t width,
                      int post, const float* row_scale, float cscale, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Two-endpoint gather per edge (no atomics):
 *   dst[e,:] = cscale * ( f(src[tail[e],:]) + f(src[head[e],:]) ),  f(v_n) = node_rcp ? (1/node_rcp[n]) * v_n : v_n
 * Replaces: torch.sparse.mm(par.abs().transpose(0,1), x_t)/2  lib/Hodge_Cheb_Conv.py:295 (=:101),
 *           and the adjoint of :294.
 * -------------------------------------------------------------------------------------------- */
int hl_endpoint_gather(const int32_t* tail, const int32_t* head, int32_t nedges,
                       const float* src, int64_t ld_src, const float* node_rcp,
                       float* dst, int64_t ld_dst, int32_t width, float cscale, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Signed boundary difference per edge, absolute value (no atomics), and its adjoint:
 *   fwd: dst[e,:]  = cscale * | src[head[e],:] - src[tail[e],:] |
 *   bwd: dsrc[n,:] = cscale * sum_{e incident to n, ascending e} s(n,e) * sgn(src[head_e]-src[tail_e]) * g[e,:],
 *        s(n,e) = +1 (n = head of e) / -1 (n = tail of e); sgn(0) = 0 like torch.abs' backward.
 * Replaces: torch.sparse.mm(par_1.transpose(0,1), x_t).abs()/2   lib/Hodge_ST_Model.py:848 (TSP readout).
 * inc_rowptr / inc_edge: node -> incident-edge CSR (hl_csr_from_coo with HL_TIE_COLUMN).
 * -------------------------------------------------------------------------------------------- */
int hl_boundary_absdiff_fwd(const int32_t* tail, const int32_t* head, int32_t nedges,
                            const float* src, int64_t ld_src, float* dst, int64_t ld_dst,
                            int32_t width, float cscale, hl_stream_t stream);
int hl_boundary_absdiff_bwd(const int32_t* inc_rowptr, const int32_t* inc_edge,
                            const int32_t* tail, const int32_t* head, int32_t nnodes,
                            const float* src, int64_t ld_src, const float* g, int64_t ld_g,
                            float* dsrc, int64_t ld_dsrc, int32_t width, float cscale, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Owner gather (adjoint of hl_segment_reduce w.r.t. src, fused with the gate's adjoint):
 *   dsrc[m,:]  = owner[m] < 0 ? 0 : g[owner[m],:] * w_m * (src_scale ? src_scale[m] : 1)
 *   dscale[m]  = owner[m] < 0 ? 0 : w_m * <g[owner[m],:], src[m,:]>        (if dscale != NULL)
 *   w_m        = owner_scale ? owner_scale[owner[m]] : 1                   (1/max(count,1))
 * Replaces: autograd of scatter_mean / global_mean_pool / `x * att` (lib/Hodge_ST_Model.py:141-150).
 * -------------------------------------------------------------------------------------------- */
int hl_owner_gather(const int32_t* owner, int32_t nsrc, const float* g, int64_t ld_g,
                    const float* owner_scale, const float* src_scale,
                    const float* src, int64_t ld_src,
                    float* dsrc, int64_t ld_dsrc, float* dscale, int32_t width, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Attention gate (lib/Hodge_Cheb_Conv.py:299-304):
 *   pre[r] = ((1-lambda) <qc[r,:], k[r,:]> + lambda <qs[r,:], k[r,:]>) / sqrt(dk)
 *   a[r]   = sigma(pre[r]),  sigma = HL_SIGMA_SIGMOID | HL_SIGMA_RELU
 * bwd: given da[r] -> dqc, dqs, dk rows.
 * -------------------------------------------------------------------------------------------- */
enum { HL_SIGMA_SIGMOID = 0, HL_SIGMA_RELU = 1 };
int hl_att_gate_fwd(const float* qc, const float* qs, const float* k, int32_t nrows, int32_t dk,
                    float lambda, int sigma, float* a, hl_stream_t stream);
int hl_att_gate_bwd(const float* qc, const float* qs, const float* k, const float* a, const float* da,
                    int32_t nrows, int32_t dk, float lambda, int sigma,
                    float* dqc, float* dqs, float* dkk, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Training-mode BatchNorm over rows + activation, writing into a column slice of a wider buffer
 * (kills the torch.cat of the dense connections).  Two deterministic stages: per-block column
 * partial sums (Welford-free: sum and sum of squares in fp32 with fp64 finalisation), then apply.
 * Replaces: gnn.BatchNorm / nn.BatchNorm1d + ReLU/LeakyReLU, lib/Hodge_ST_Model.py:580-586,
 * lib/Hodge_Cheb_Conv.py:278-282, and torch.cat lib/Hodge_ST_Model.py:632-633.
 *   y = act( (x - mean) * rsqrt(var_biased + eps) * gamma + beta ),  act: slope 0 = ReLU, 1 = identity
 * stats[0:F] = mean, stats[F:2F] = biased var (both fp32).  If running_mean/running_var are given they are
 * updated in the same launch as nn.BatchNorm1d does (momentum, unbiased variance).
 * `nvalid` (device int32, nullable): only rows [0, *nvalid) enter the statistics; rows beyond are
 * padding ("ghost" rows of a fixed-capacity batch replayed from a CUDA graph): y and dx are written
 * as exact zeros there, so they contribute nothing to any later reduction or weight gradient.
 * -------------------------------------------------------------------------------------------- */
size_t hl_bn_workspace(int32_t nrows, int32_t width);
int hl_bn_act_fwd(const float* x, int64_t ld_x, int32_t nrows, int32_t width,
                  const float* gamma, const float* beta, float eps, float slope,
                  float* y, int64_t ld_y, float* stats, const int32_t* nvalid /* device, nullable */,
                  float* running_mean /* nullable */, float* running_var, float momentum,
                  int64_t* num_batches_tracked /* device, nullable: incremented by one in the same launch */,
                  void* workspace, size_t workspace_bytes, hl_stream_t stream);
/* hl_bn_act_fwd without its statistics pass: `bn_part` ([ceil(nrows/32)][2][width]: mean | M2 of every 32-row block) was
 * written by the epilogue of the GEMM that produced x (hl_gemm2_bn_tf32x3, the Linear -> BatchNorm pairs of
 * lib/Hodge_Cheb_Conv.py:277-288 and lib/Hodge_ST_Model.py:578-590); merged with Chan's formula in fp64, then applied. */
int hl_bn_act_fwd_tiles(const float* x, int64_t ld_x, int32_t nrows, int32_t width,
                        const float* gamma, const float* beta, float eps, float slope,
                        float* y, int64_t ld_y, float* stats, const int32_t* nvalid,
                        float* running_mean, float* running_var, float momentum,
                        int64_t* num_batches_tracked, const float* bn_part, hl_stream_t stream);
int hl_bn_act_bwd(const float* x, int64_t ld_x, const float* y, int64_t ld_y,
                  const float* dy, int64_t ld_dy, const float* dy2 /* nullable: a second gradient piece, added on the fly */, int64_t ld_dy2,
                  int32_t nrows, int32_t width,
                  const float* gamma, const float* stats, float eps, float slope,
                  float* dx, int64_t ld_dx, float* dgamma, float* dbeta,
                  int accumulate_param_grads /* dgamma / dbeta: 0 = overwrite, 1 = += (fused gradient accumulation) */,
                  const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream);

/* The same BatchNorm as separate phases, for statistics synchronised across data-parallel ranks (SURVEY section 8e,
 * caveat 1: the reference trains the whole batch on one GPU, so matching a single-GPU run of the GLOBAL batch needs
 * batch statistics over all ranks).  Forward: hl_bn_stats (per-rank mean | biased variance) -> the caller combines
 * the ranks' statistics -> hl_bn_apply with the global ones.  Backward: hl_bn_bwd_sums (per-rank column sums of
 * dz and dz * xhat, dz = dy * act'(y); these are also this rank's dbeta / dgamma) -> the caller sums them over the
 * ranks -> hl_bn_bwd_apply with the global sums and *inv_count = 1 / (rows over all ranks) (device scalar; NULL =
 * this rank's own row count). */
int hl_bn_stats(const float* x, int64_t ld_x, int32_t nrows, int32_t width, float* stats /* [2*width] */,
                const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream);
int hl_bn_apply(const float* x, int64_t ld_x, int32_t nrows, int32_t width, const float* gamma, const float* beta,
                const float* stats, float eps, float slope, float* y, int64_t ld_y, const int32_t* nvalid,
                hl_stream_t stream);
int hl_bn_bwd_sums(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy, int64_t ld_dy,
                   int32_t nrows, int32_t width, const float* stats, float eps, float slope, float* sums /* [2*width] */,
                   const int32_t* nvalid, void* workspace, size_t workspace_bytes, hl_stream_t stream);
int hl_bn_bwd_apply(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy, int64_t ld_dy,
                    int32_t nrows, int32_t width, const float* gamma, const float* stats, const float* sums,
                    const float* inv_count /* device, nullable */, float eps, float slope, float* dx, int64_t ld_dx,
                    const int32_t* nvalid, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Adam over flat buffers: one streaming pass over parameters, gradients and both moments (all [n] fp32, 16-byte aligned).
 * Replaces: torch.optim.Adam(model.parameters(), lr, weight_decay).step() of the training scripts
 * (main_zinc_HL_HGCNN_dense_int3_pyr.py:213, :150-160) -- same update rule (L2 weight decay added to the gradient,
 * bias-corrected moments, eps outside the root); `grad_scale` multiplies the gradient first (1 / world_size of the
 * data-parallel average).  state[3] (device): {step, 1 / (1 - beta1^step), 1 / sqrt(1 - beta2^step)}; the call advances it,
 * so the optimizer step is capturable into a CUDA graph.
 * -------------------------------------------------------------------------------------------- */
int hl_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state,
                 float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Eigenvector positional encodings of every graph of a mini-batch.
 * Replaces: `eig_pe(L, k)` = scipy.linalg.eigh of the dense normalised Laplacian + argsort + columns 1 .. k-1
 * (lib/Hodge_Dataset.py:97-112; callers :457-458 ZINC, :586-587 peptides, :846-847 CIFAR10SP) and the dense `eigh`
 * behind it, per graph on the CPU.  The diagonal blocks of a block-diagonal CSR operator (rows seg_ptr[g] ..
 * seg_ptr[g+1] of `rowptr / colidx / vals`, e.g. L0 or L1 from hl_laplacian_fill or hl_csr_from_coo) are expanded
 * into dense fp64 scratch and diagonalised by one CTA each (cyclic Jacobi, round-robin pair order).
 *   evals[R]          eigenvalues of every block in ascending order (R = seg_ptr[n_graphs])
 *   pe[R, k-1]        eigenvectors of rank 1 .. k-1 (rank 0, the constant vector of a connected graph, is skipped like
 *                     the reference's `eig_vecs[:, 1:k]`); blocks with fewer than k rows are zero-padded, which is what
 *                     the datasets' `get` does (:430-431); sign: largest-magnitude component positive (the reference
 *                     inherits LAPACK's arbitrary sign and re-draws it at random per sample, :429-439)
 *   vecs_all          optional (nullable): all eigenvectors, [n_g, n_g] row-major per block at mat_ptr[g]
 *   sweeps            optional (nullable): Jacobi sweeps used per block
 * mat_ptr[g] = sum_{h<g} n_h^2 (int64, device), total_matrix_elements = mat_ptr[n_graphs] (host), max_n = max n_g.
 * -------------------------------------------------------------------------------------------- */
size_t hl_eig_pe_workspace(int64_t total_matrix_elements);
int hl_eig_pe(const int32_t* seg_ptr, int32_t n_graphs, int32_t max_n, const int64_t* mat_ptr,
              int64_t total_matrix_elements, const int32_t* rowptr, const int32_t* colidx, const float* vals,
              int32_t k, float* evals, float* pe, int64_t ld_pe, float* vecs_all /* nullable */,
              int32_t* sweeps /* nullable */, int32_t max_sweeps, void* workspace, size_t workspace_bytes,
              hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Greedy heavy-edge matching for multi-level graph coarsening, one warp per graph of the mini-batch.
 * Nodes are visited in id order; an unmatched node u pairs with its unmatched neighbour of largest edge_weight
 * (NULL = all ones; ties: first in ascending incident-edge order); cluster[u] = cluster[partner] = u.
 * Replaces: torch_cluster.graclus_cluster in MLGC / MLGC_weighted (lib/Hodge_Dataset.py:254-255, :309-311) --
 * the library routine is randomised; this is the deterministic variant the CPU oracle pins.
 * node_ptr[G+1]: node range of every graph; inc_rowptr/inc_edge: node -> incident-edge CSR; tail/head per edge.
 * -------------------------------------------------------------------------------------------- */
int hl_greedy_matching(const int32_t* node_ptr, int32_t n_graphs, const int32_t* inc_rowptr, const int32_t* inc_edge,
                       const int32_t* tail, const int32_t* head, const float* edge_weight /* nullable */,
                       int32_t* cluster /* out [N] */, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * fp32-accurate dense transform on the tcgen05 tensor cores (3xTF32 split, fp32 TMEM accumulator):
 *   hl_gemm_tf32x3 : C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) , or C += ... when accumulate != 0
 * A row-major [M,K] (pitch lda), B = the layer weight [N,K] pre-split by hl_tf32_split into
 * hi (tf32-exact) and lo (remainder) copies [N,K] (pitch ldb).  Replaces the cuBLAS SGEMM behind
 * `lins[k](T_k)` / nn.Linear (lib/Hodge_Cheb_Conv.py:487,497,509, :277-288) and its data gradient
 * (B = W^T, hl_tf32_split(transpose=1)).  Requirements: lda, ldb multiples of 4, 16-byte aligned
 * bases, N a multiple of 16 (one CTA covers 128 rows x min(N,256) columns).  Returns 1 (not an
 * error) when the shape is unsupported so the caller can use a library GEMM.
 * -------------------------------------------------------------------------------------------- */
int hl_tf32_split(const float* src, int64_t ld_src, int32_t rows, int32_t cols, int transpose,
                  float* hi, float* lo, int64_t ld_out, hl_stream_t stream);
/* The same split for a table of weight views in ONE launch (all weights of a model at the top of a training step,
 * off the critical chain of the GEMMs that consume them).  `table`: DEVICE array of n_entries descriptors;
 * max_elements = max rows*cols over the table (sizes the grid). */
typedef struct hl_split_desc {
  const float* src;   /* [rows, cols], row pitch ld_src */
  float* hi;          /* [rows, cols] (or [cols, rows] when transpose), row pitch ld_out */
  float* lo;
  int64_t ld_src;
  int64_t ld_out;
  int32_t rows;
  int32_t cols;
  int32_t transpose;
  int32_t reserved;
} hl_split_desc;
int hl_tf32_split_batch(const hl_split_desc* table, int32_t n_entries, int64_t max_elements, hl_stream_t stream);
int hl_gemm_tf32x3(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb,
                   int32_t M, int32_t N, int32_t K, const float* bias, float* C, int64_t ldc,
                   int accumulate, hl_stream_t stream);
/* C = [A | A2] * B^T in one launch: the contraction continues over a second row-major operand (the next
 * polynomial order T_1 after T_0 = x, or [transferred | own] of the NodeEdgeInt MLP), so the partial sum never
 * leaves TMEM.  B is packed [N, pad32(K) + K2] (block 2 starts at column pad32(K); padding must be zero). */
int hl_gemm2_tf32x3(const float* A, int64_t lda, int32_t K, const float* A2, int64_t lda2, int32_t K2,
                    const float* Bhi, const float* Blo, int64_t ldb, int32_t M, int32_t N,
                    const float* bias, float* C, int64_t ldc, int accumulate, hl_stream_t stream);
/* ... with the BatchNorm statistics of the FINAL output values from the epilogue: bn_part [hl_gemm_bn_part_floats(M, N)] gets
 * (mean | M2) of every 32-row block over the rows below *bn_nvalid (NULL = M); pass it on the last launch that touches C. */
size_t hl_gemm_bn_part_floats(int32_t M, int32_t N);
int hl_gemm2_bn_tf32x3(const float* A, int64_t lda, int32_t K, const float* A2, int64_t lda2, int32_t K2,
                       const float* Bhi, const float* Blo, int64_t ldb, int32_t M, int32_t N,
                       const float* bias, float* C, int64_t ldc, int accumulate, float* bn_part /* nullable */,
                       const int32_t* bn_nvalid /* nullable */, hl_stream_t stream);
/* Weight gradient on the same tensor-core path: dw[fo,fi] (=|+=) g[R,fo]^T x[R,fi].  Both operands are
 * consumed MN-major straight from their row-major storage ({32 x 32} TMA boxes, no transposes), split into
 * hi/lo inside the kernel, the R rows are divided over CTAs and the partial tiles are summed in a fixed
 * order (deterministic).  Needs fo % 4 == 0, fi % 32 == 0, R >= 512; returns 1 otherwise (use hl_wgrad). */
size_t hl_wgrad_tf32x3_workspace(int32_t nrows, int32_t fo, int32_t fi);
int hl_wgrad_tf32x3(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo, int32_t fi,
                    float* dw, int64_t ld_dw, int accumulate, void* workspace, size_t workspace_bytes,
                    hl_stream_t stream);
/* The same, plus dbias[fo] (=|+=) the column sums of g (the bias gradient of the Linear / conv whose weight gradient
 * this is) folded into the same two launches.  Returns 2 when dW was computed but dbias was not (use hl_colsum). */
int hl_wgrad_bias_tf32x3(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo, int32_t fi,
                         float* dw, int64_t ld_dw, int accumulate, float* dbias /* nullable */, int accumulate_bias,
                         void* workspace, size_t workspace_bytes, hl_stream_t stream);
/* Two activations of the same shape sharing g: dW1 (=|+=) g^T x1, dW2 (=|+=) g^T x2 from ONE tensor-core launch and one
 * reduce (the two halves of the first NodeEdgeInt Linear, lib/Hodge_Cheb_Conv.py:307-308; consecutive orders of a conv,
 * :497,509).  Returns 1 when the shape is unsupported. */
size_t hl_wgrad2_tf32x3_workspace(int32_t nrows, int32_t fo, int32_t fi);
int hl_wgrad2_bias_tf32x3(const float* g, int64_t ld_g, const float* x1, int64_t ld_x1, const float* x2, int64_t ld_x2,
                          int32_t nrows, int32_t fo, int32_t fi, float* dw1, int64_t ld_dw1, float* dw2, int64_t ld_dw2,
                          int accumulate, float* dbias /* nullable */, int accumulate_bias, void* workspace,
                          size_t workspace_bytes, hl_stream_t stream);

/* Deferred split reduce: hl_wgrad_deferred_tf32x3 launches only the tensor-core pass of a weight gradient (one activation:
 * x2 = dw2 = NULL, or two that share g) and describes the pending reduce in *desc; `workspace` holds the partial planes and
 * must live until hl_wgrad_reduce_batch has summed them.  A training step collects the descriptors of all its weight
 * gradients (lib/Hodge_Cheb_Conv.py:487-510 and :277-288 backward, dozens per step) and reduces them in ONE launch before
 * the all-reduce; two descriptors of one batch must not target the same elements.  `descs` is a host array. */
typedef struct hl_wgrad_reduce_desc {
  const float* partial;     /* [splits] planes of fo x fi, split_stride elements apart */
  const float* cs_partial;  /* [2 splits][fo] column sums of g (bias gradient), used when dbias != NULL */
  float* dw;                /* destination of columns [0, fi_first) (all columns when dw2 == NULL) */
  float* dw2;               /* destination of columns [fi_first, fi), or NULL */
  float* dbias;             /* or NULL */
  int64_t split_stride;
  int64_t ld_dw;
  int64_t ld_dw2;
  int32_t splits;
  int32_t fo;
  int32_t fi;               /* columns of a partial plane */
  int32_t fi_first;
  int32_t accumulate;
  int32_t accumulate_bias;
  int32_t block_start;      /* filled by hl_wgrad_reduce_batch */
  int32_t reserved;
} hl_wgrad_reduce_desc;
int hl_wgrad_deferred_tf32x3(const float* g, int64_t ld_g, const float* x1, int64_t ld_x1, const float* x2 /* nullable */,
                             int64_t ld_x2, int32_t nrows, int32_t fo, int32_t fi, float* dw1, int64_t ld_dw1,
                             float* dw2 /* nullable */, int64_t ld_dw2, int accumulate, float* dbias /* nullable */,
                             int accumulate_bias, void* workspace, size_t workspace_bytes, hl_wgrad_reduce_desc* desc,
                             hl_stream_t stream);
int hl_wgrad_reduce_batch(const hl_wgrad_reduce_desc* descs, int32_t n, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Weight and bias gradients of the dense layers as deterministic split-row reductions.
 *   hl_wgrad : dw[fo,fi] (=|+=) g[R,fo]^T x[R,fi]        hl_colsum : out[f] (=|+=) sum_r g[r,f]
 * Replaces: autograd of `lins[k](T_k)` / nn.Linear (`lib/Hodge_Cheb_Conv.py:487,497,509`, `:277-288`):
 * cuBLAS walks all R rows with one CTA per 64x64 output tile; here rows are split over CTAs, partial
 * tiles are summed in a fixed order (no atomics).  fp32 FMA.
 * -------------------------------------------------------------------------------------------- */
size_t hl_wgrad_workspace(int32_t nrows, int32_t fo, int32_t fi);
int hl_wgrad(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo, int32_t fi,
             float* dw, int64_t ld_dw, int accumulate, void* workspace, size_t workspace_bytes, hl_stream_t stream);
size_t hl_colsum_workspace(int32_t nrows, int32_t width);
int hl_colsum(const float* g, int64_t ld_g, int32_t nrows, int32_t width, float* out, int accumulate,
              void* workspace, size_t workspace_bytes, hl_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Simplex-graph construction for a whole mini-batch (block-diagonal), on the GPU.
 * Replaces: Dataset.process lib/Hodge_Dataset.py:447-456,467-468 (to_undirected -> i<j ->
 * dense B1 -> B1 B1^T -> eigh -> 2L/lambda_max -> dense_to_sparse), the same tail of MLGC
 * (:276-288) and the block-diagonal collation of PairData.__inc__ (:40-48): the caller passes the
 * directed edges of ALL graphs with global (graph-contiguous) node ids.
 *
 * step 1  hl_build_edges: directed (any order, duplicates and both directions allowed) -> unique
 *         undirected i<j edges in lexicographic order (= to_undirected + mask, :447-450).
 *         tail/head/attr_out need room for n_directed entries; *n_edges_out (device) = count.
 *         attr (optional int64 per directed edge) is min-reduced over duplicates (reduce='min').
 * step 2  node -> incident-edge CSR: hl_csr_from_coo(row = [tail | head], col = [e | e],
 *         tie = HL_TIE_COLUMN)  (rowptr diff = node degree).
 * step 3  hl_lambda_max: lambda_max(B1 B1^T) per graph, Lanczos with full re-orthogonalisation in
 *         fp64, one CTA per graph, at most min(n_g, max_steps) steps (exact when max_steps >=
 *         n_g); last_change[g] = |theta_m - theta_{m-1}| of the final step (convergence evidence).
 * step 4  hl_laplacian_rowptr (row counts + scan; nnz = rowptr[last]) then hl_laplacian_fill:
 *         CSR of L0 = 2 B1 B1^T / lmax (row i: ascending neighbours, diagonal = degree; isolated
 *         node = empty row) and L1 = 2 B1^T B1 / lmax (row e: ascending union of the edges
 *         incident to its endpoints; +1 if both edges leave / both enter the shared node, else -1;
 *         diagonal 2), values = fp32(2*m)/lmax[graph] as the reference computes them.  Column
 *         order equals the reference's row-major dense_to_sparse COO, so (rowptr, col, val) is at
 *         once the forward and (L symmetric) the transposed operator for hl_poly_*.
 * -------------------------------------------------------------------------------------------- */
size_t hl_build_edges_workspace(int64_t n_directed);
int hl_build_edges(const int64_t* src, const int64_t* dst, int64_t n_directed, int64_t n_nodes,
                   const int64_t* attr, int32_t* tail, int32_t* head, int64_t* attr_out,
                   int32_t* n_edges_out, void* workspace, size_t workspace_bytes, hl_stream_t stream);
size_t hl_lambda_max_workspace(int32_t n_graphs, int32_t max_nodes, int32_t max_steps);
int hl_lambda_max(const int32_t* node_ptr /* [G+1] */, int32_t n_graphs, int32_t max_nodes,
                  const int32_t* inc_rowptr, const int32_t* inc_edge,
                  const int32_t* tail, const int32_t* head, int32_t max_steps,
                  float* lambda_max, float* last_change, void* workspace, size_t workspace_bytes,
                  hl_stream_t stream);
size_t hl_laplacian_rowptr_workspace(int32_t n_edges, int32_t n_nodes);
int hl_laplacian_rowptr(const int32_t* tail, const int32_t* head, int32_t n_edges, int32_t n_nodes,
                        const int32_t* inc_rowptr, int32_t* l0_rowptr /* [N+1] */, int32_t* l1_rowptr /* [E+1] */,
                        void* workspace, size_t workspace_bytes, hl_stream_t stream);
int hl_laplacian_fill(const int32_t* tail, const int32_t* head, int32_t n_edges, int32_t n_nodes,
                      const int32_t* inc_rowptr, const int32_t* inc_edge,
                      const int32_t* node_graph, const float* lambda_max /* per graph */,
                      const int32_t* l0_rowptr, int32_t* l0_col, float* l0_val,
                      const int32_t* l1_rowptr, int32_t* l1_col, float* l1_val, hl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HLHGAT_H_ */
