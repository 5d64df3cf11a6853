// GPU simplex-graph construction for a whole block-diagonal mini-batch:
// directed edge list -> unique undirected i<j edges in lexicographic order -> CSR of
// L0 = 2 B1 B1^T / lmax and L1 = 2 B1^T B1 / lmax with ascending columns (the order the reference's
// dense_to_sparse emits), never materialising the dense N x E boundary matrix, plus the per-graph
// lambda_max(B1 B1^T) by Lanczos with full re-orthogonalisation in fp64.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace hl {

constexpr uint64_t kDropKey = ~0ull;

// ---------------------------------------------------------------------------------------------
// step 1: unique undirected edges
// ---------------------------------------------------------------------------------------------
__global__ void edge_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t m,
                                 int64_t n_nodes, uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = src[i], b = dst[i];
    const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    const bool ok = lo != hi && lo >= 0 && hi < n_nodes;       // self loops never pass the i<j mask
    keys[i] = ok ? (uint64_t)lo * (uint64_t)n_nodes + (uint64_t)hi : kDropKey;
    idx[i] = (int32_t)i;
  }
}

__global__ void edge_flags_kernel(const uint64_t* __restrict__ keys, int64_t m, int32_t* __restrict__ flags) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[p];
    flags[p] = (k != kDropKey && (p == 0 || keys[p - 1] != k)) ? 1 : 0;
  }
}

__global__ void edge_emit_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx,
                                 const int32_t* __restrict__ flags, const int32_t* __restrict__ pos, int64_t m,
                                 int64_t n_nodes, const int64_t* __restrict__ attr, int32_t* __restrict__ tail,
                                 int32_t* __restrict__ head, int64_t* __restrict__ attr_out,
                                 int32_t* __restrict__ n_edges_out) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (int64_t)gridDim.x * blockDim.x) {
    if (p == m - 1) *n_edges_out = pos[p] + flags[p];
    if (!flags[p]) continue;
    const uint64_t k = keys[p];
    const int32_t e = pos[p];
    tail[e] = (int32_t)(k / (uint64_t)n_nodes);
    head[e] = (int32_t)(k % (uint64_t)n_nodes);
    if (attr && attr_out) {                                    // to_undirected(reduce='min') over duplicates
      int64_t best = attr[idx[p]];
      for (int64_t q = p + 1; q < m && keys[q] == k; ++q) {
        const int64_t v = attr[idx[q]];
        best = v < best ? v : best;
      }
      attr_out[e] = best;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// step 3: Laplacian CSR from the incidence lists
// ---------------------------------------------------------------------------------------------
__global__ void laplacian_counts_kernel(const int32_t* __restrict__ tail, const int32_t* __restrict__ head,
                                        int32_t n_edges, int32_t n_nodes, const int32_t* __restrict__ inc_rowptr,
                                        int32_t* __restrict__ c0, int32_t* __restrict__ c1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes) {
    const int d = inc_rowptr[i + 1] - inc_rowptr[i];
    c0[i] = d > 0 ? d + 1 : 0;                                 // isolated node: all-zero row, no diagonal entry
  }
  if (i < n_edges) {
    const int t = tail[i], h = head[i];
    c1[i] = (inc_rowptr[t + 1] - inc_rowptr[t]) + (inc_rowptr[h + 1] - inc_rowptr[h]) - 1;
  }
  if (i == 0) { c0[n_nodes] = 0; c1[n_edges] = 0; }
}

__device__ __forceinline__ float scaled(float integer_entry, float lmax) {
  return __fdiv_rn(__fmul_rn(2.f, integer_entry), lmax);       // (2*M)/maxeig, as lib/Hodge_Dataset.py:455-456
}

__global__ void laplacian_fill_kernel(const int32_t* __restrict__ tail, const int32_t* __restrict__ head,
                                      int32_t n_edges, int32_t n_nodes, const int32_t* __restrict__ inc_rowptr,
                                      const int32_t* __restrict__ inc_edge, const int32_t* __restrict__ node_graph,
                                      const float* __restrict__ lambda_max,
                                      const int32_t* __restrict__ r0, int32_t* __restrict__ col0, float* __restrict__ val0,
                                      const int32_t* __restrict__ r1, int32_t* __restrict__ col1, float* __restrict__ val1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes) {
    const int s = inc_rowptr[i], e = inc_rowptr[i + 1], d = e - s;
    if (d > 0) {
      const float lm = lambda_max[node_graph[i]];
      int w = r0[i];
      bool diag_done = false;
      for (int p = s; p < e; ++p) {                            // incident edges ascending => neighbours ascending
        const int ed = inc_edge[p];
        const int other = tail[ed] == i ? head[ed] : tail[ed];
        if (!diag_done && other > i) {
          col0[w] = i; val0[w] = scaled((float)d, lm); ++w;
          diag_done = true;
        }
        col0[w] = other; val0[w] = scaled(-1.f, lm); ++w;
      }
      if (!diag_done) { col0[w] = i; val0[w] = scaled((float)d, lm); }
    }
  }
  if (i < n_edges) {
    const int t = tail[i], h = head[i];
    const float lm = lambda_max[node_graph[t]];
    int pa = inc_rowptr[t], ea = inc_rowptr[t + 1], pb = inc_rowptr[h], eb = inc_rowptr[h + 1];
    int w = r1[i];
    while (pa < ea || pb < eb) {                               // merge of two ascending edge-id lists
      const int fa = pa < ea ? inc_edge[pa] : 0x7fffffff;
      const int fb = pb < eb ? inc_edge[pb] : 0x7fffffff;
      int f, shared;
      if (fa <= fb) { f = fa; shared = t; ++pa; if (fb == fa) ++pb; }
      else { f = fb; shared = h; ++pb; }
      float m;
      if (f == i) m = 2.f;
      else m = ((tail[f] == shared) == (tail[i] == shared)) ? 1.f : -1.f;   // same orientation at the shared node
      col1[w] = f; val1[w] = scaled(m, lm); ++w;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// lambda_max(D - A) per graph: Lanczos, full re-orthogonalisation, fp64, one CTA per graph
// ---------------------------------------------------------------------------------------------
constexpr int kLzThreads = 128;

__device__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kLzThreads / 32; ++w) t += red[w];
  return t;
}

// largest eigenvalue of the symmetric tridiagonal (alpha[0..m), beta[0..m-1)) by Sturm bisection
__device__ double tridiag_max_eig(const double* alpha, const double* beta, int m) {
  double lo = alpha[0], hi = alpha[0];
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i + 1 < m ? fabs(beta[i]) : 0.0);
    lo = fmin(lo, alpha[i] - r);
    hi = fmax(hi, alpha[i] + r);
  }
  for (int it = 0; it < 100 && hi - lo > 1e-15 * fmax(fabs(hi), 1.0); ++it) {
    const double mid = 0.5 * (lo + hi);
    int below = 0;                                             // eigenvalues < mid
    double q = 1.0;
    for (int i = 0; i < m; ++i) {
      const double b2 = i > 0 ? beta[i - 1] * beta[i - 1] : 0.0;
      q = alpha[i] - mid - (i > 0 ? b2 / q : 0.0);
      if (q == 0.0) q = 1e-300;
      if (q < 0.0) ++below;
    }
    if (below == m) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

__global__ void __launch_bounds__(kLzThreads)
lambda_max_kernel(const int32_t* __restrict__ node_ptr, int32_t max_nodes, const int32_t* __restrict__ inc_rowptr,
                  const int32_t* __restrict__ inc_edge, const int32_t* __restrict__ tail, const int32_t* __restrict__ head,
                  int32_t max_steps, float* __restrict__ lambda_max, float* __restrict__ last_change,
                  double* __restrict__ work) {
  __shared__ double red[kLzThreads / 32];
  __shared__ double sh_theta, sh_prev;
  __shared__ int sh_stop;
  const int g = blockIdx.x;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
  if (n <= 0) { if (threadIdx.x == 0) { lambda_max[g] = 0.f; if (last_change) last_change[g] = 0.f; } return; }
  const int m_max = min(n, max_steps);
  // workspace per graph: Q [m_max+1][n], w[n], alpha[m_max], beta[m_max]
  double* Q = work + (size_t)g * ((size_t)(max_steps + 2) * max_nodes + 2 * (size_t)max_steps);
  double* wv = Q + (size_t)(max_steps + 1) * max_nodes;
  double* alpha = wv + max_nodes;
  double* beta = alpha + max_steps;

  // deterministic start vector with a component along every coordinate
  double loc = 0.0;
  for (int i = threadIdx.x; i < n; i += kLzThreads) {
    unsigned h = (unsigned)(i + 1) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const double v = 0.5 + (double)(h & 0xffff) / 65536.0;
    Q[i] = v;
    loc += v * v;
  }
  double nrm = sqrt(block_sum(loc, red));
  for (int i = threadIdx.x; i < n; i += kLzThreads) Q[i] /= nrm;
  if (threadIdx.x == 0) { sh_theta = 0.0; sh_prev = -1.0; sh_stop = 0; }
  __syncthreads();

  int m = 0;
  for (int j = 0; j < m_max; ++j) {
    const double* q = Q + (size_t)j * max_nodes;
    // w = (D - A) q
    loc = 0.0;
    for (int i = threadIdx.x; i < n; i += kLzThreads) {
      const int node = n0 + i;
      const int s = inc_rowptr[node], e = inc_rowptr[node + 1];
      double acc = (double)(e - s) * q[i];
      for (int p = s; p < e; ++p) {
        const int ed = inc_edge[p];
        const int other = tail[ed] == node ? head[ed] : tail[ed];
        acc -= q[other - n0];
      }
      wv[i] = acc;
      loc += acc * q[i];
    }
    const double a = block_sum(loc, red);
    if (threadIdx.x == 0) alpha[j] = a;
    // full re-orthogonalisation against q_0..q_j (twice is enough)
    for (int pass = 0; pass < 2; ++pass) {
      for (int k = 0; k <= j; ++k) {
        const double* qk = Q + (size_t)k * max_nodes;
        loc = 0.0;
        for (int i = threadIdx.x; i < n; i += kLzThreads) loc += wv[i] * qk[i];
        const double c = block_sum(loc, red);
        for (int i = threadIdx.x; i < n; i += kLzThreads) wv[i] -= c * qk[i];
        __syncthreads();
      }
    }
    loc = 0.0;
    for (int i = threadIdx.x; i < n; i += kLzThreads) loc += wv[i] * wv[i];
    const double b = sqrt(block_sum(loc, red));
    m = j + 1;
    if (threadIdx.x == 0) {
      beta[j] = b;
      const double th = tridiag_max_eig(alpha, beta, m);
      sh_prev = sh_theta;
      sh_theta = th;
      // invariant subspace reached, or the Ritz value has stopped moving
      sh_stop = (b <= 1e-12 * fmax(fabs(th), 1.0)) || (m >= 4 && fabs(th - sh_prev) <= 1e-13 * fabs(th));
    }
    __syncthreads();
    if (sh_stop || j + 1 == m_max) break;
    double* qn = Q + (size_t)(j + 1) * max_nodes;
    for (int i = threadIdx.x; i < n; i += kLzThreads) qn[i] = wv[i] / b;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    lambda_max[g] = (float)sh_theta;
    if (last_change) last_change[g] = (float)fabs(sh_theta - sh_prev);
  }
}

static int grid1d(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static size_t sort_bytes(int64_t m) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, m);
  return bytes;
}

static size_t scan_bytes(int64_t m) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, m);
  return bytes;
}

}  // namespace hl

extern "C" size_t hl_build_edges_workspace(int64_t n_directed) {
  if (n_directed < 0) return 0;
  const size_t m = (size_t)(n_directed > 0 ? n_directed : 1);
  const size_t tmp = hl::sort_bytes(n_directed) > hl::scan_bytes(n_directed) ? hl::sort_bytes(n_directed) : hl::scan_bytes(n_directed);
  return hl::align_up(m * 8, 256) * 2 + hl::align_up(m * 4, 256) * 4 + hl::align_up(tmp, 256) + 256;
}

extern "C" int hl_build_edges(const int64_t* src, const int64_t* dst, int64_t n_directed, int64_t n_nodes,
                              const int64_t* attr, int32_t* tail, int32_t* head, int64_t* attr_out,
                              int32_t* n_edges_out, void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (n_directed < 0 || n_nodes < 0 || n_directed > 0x7fffffffLL || n_nodes > 0x7fffffffLL || !n_edges_out) return HL_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (n_directed == 0) {
    HL_CUDA_CHECK(cudaMemsetAsync(n_edges_out, 0, sizeof(int32_t), st));
    return HL_OK;
  }
  if (!src || !dst || !tail || !head) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_build_edges_workspace(n_directed)) return HL_ERR_WORKSPACE;
  const size_t m = (size_t)n_directed;
  char* w = reinterpret_cast<char*>(workspace);
  uint64_t* k_in = reinterpret_cast<uint64_t*>(w);  w += align_up(m * 8, 256);
  uint64_t* k_out = reinterpret_cast<uint64_t*>(w); w += align_up(m * 8, 256);
  int32_t* i_in = reinterpret_cast<int32_t*>(w);    w += align_up(m * 4, 256);
  int32_t* i_out = reinterpret_cast<int32_t*>(w);   w += align_up(m * 4, 256);
  int32_t* flags = reinterpret_cast<int32_t*>(w);   w += align_up(m * 4, 256);
  int32_t* pos = reinterpret_cast<int32_t*>(w);     w += align_up(m * 4, 256);
  void* tmp = w;
  size_t sb = sort_bytes(n_directed), cb = scan_bytes(n_directed);
  const int g = grid1d(n_directed, 256);
  edge_keys_kernel<<<g, 256, 0, st>>>(src, dst, n_directed, n_nodes, k_in, i_in);
  HL_LAUNCH_CHECK("edge_keys_kernel");
  HL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, sb, k_in, k_out, i_in, i_out, n_directed, 0, 64, st));
  edge_flags_kernel<<<g, 256, 0, st>>>(k_out, n_directed, flags);
  HL_LAUNCH_CHECK("edge_flags_kernel");
  HL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, cb, flags, pos, n_directed, st));
  edge_emit_kernel<<<g, 256, 0, st>>>(k_out, i_out, flags, pos, n_directed, n_nodes, attr, tail, head, attr_out, n_edges_out);
  HL_LAUNCH_CHECK("edge_emit_kernel");
  return HL_OK;
}

extern "C" size_t hl_laplacian_rowptr_workspace(int32_t n_edges, int32_t n_nodes) {
  const int64_t m = (int64_t)(n_edges > n_nodes ? n_edges : n_nodes) + 1;
  return hl::align_up(hl::scan_bytes(m), 256) + 256;
}

extern "C" int hl_laplacian_rowptr(const int32_t* tail, const int32_t* head, int32_t n_edges, int32_t n_nodes,
                                   const int32_t* inc_rowptr, int32_t* l0_rowptr, int32_t* l1_rowptr,
                                   void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (n_edges < 0 || n_nodes < 0 || !inc_rowptr || !l0_rowptr || !l1_rowptr || (n_edges > 0 && (!tail || !head))) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_laplacian_rowptr_workspace(n_edges, n_nodes)) return HL_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int n = (n_edges > n_nodes ? n_edges : n_nodes);
  laplacian_counts_kernel<<<(n + 256) / 256, 256, 0, st>>>(tail, head, n_edges, n_nodes, inc_rowptr, l0_rowptr, l1_rowptr);
  HL_LAUNCH_CHECK("laplacian_counts_kernel");
  size_t b0 = scan_bytes(n_nodes + 1), b1 = scan_bytes(n_edges + 1);
  HL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(workspace, b0, l0_rowptr, l0_rowptr, n_nodes + 1, st));
  HL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(workspace, b1, l1_rowptr, l1_rowptr, n_edges + 1, st));
  return HL_OK;
}

extern "C" int hl_laplacian_fill(const int32_t* tail, const int32_t* head, int32_t n_edges, int32_t n_nodes,
                                 const int32_t* inc_rowptr, const int32_t* inc_edge,
                                 const int32_t* node_graph, const float* lambda_max,
                                 const int32_t* l0_rowptr, int32_t* l0_col, float* l0_val,
                                 const int32_t* l1_rowptr, int32_t* l1_col, float* l1_val, hl_stream_t stream) {
  using namespace hl;
  if (n_edges < 0 || n_nodes < 0) return HL_ERR_INVALID;
  if (n_edges == 0 || n_nodes == 0) return HL_OK;
  if (!tail || !head || !inc_rowptr || !inc_edge || !node_graph || !lambda_max || !l0_rowptr || !l0_col || !l0_val ||
      !l1_rowptr || !l1_col || !l1_val)
    return HL_ERR_INVALID;
  const int n = (n_edges > n_nodes ? n_edges : n_nodes);
  laplacian_fill_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(tail, head, n_edges, n_nodes, inc_rowptr, inc_edge,
                                                                         node_graph, lambda_max, l0_rowptr, l0_col, l0_val,
                                                                         l1_rowptr, l1_col, l1_val);
  HL_LAUNCH_CHECK("laplacian_fill_kernel");
  return HL_OK;
}

extern "C" size_t hl_lambda_max_workspace(int32_t n_graphs, int32_t max_nodes, int32_t max_steps) {
  if (n_graphs < 0 || max_nodes < 0 || max_steps < 0) return 0;
  return ((size_t)(max_steps + 2) * (size_t)max_nodes + 2 * (size_t)max_steps) * sizeof(double) * (size_t)(n_graphs > 0 ? n_graphs : 1);
}

extern "C" int hl_lambda_max(const int32_t* node_ptr, int32_t n_graphs, int32_t max_nodes,
                             const int32_t* inc_rowptr, const int32_t* inc_edge,
                             const int32_t* tail, const int32_t* head, int32_t max_steps,
                             float* lambda_max, float* last_change, void* workspace, size_t workspace_bytes,
                             hl_stream_t stream) {
  using namespace hl;
  if (n_graphs < 0 || max_nodes < 1 || max_steps < 1) return HL_ERR_INVALID;
  if (n_graphs == 0) return HL_OK;
  if (!node_ptr || !inc_rowptr || !lambda_max) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_lambda_max_workspace(n_graphs, max_nodes, max_steps)) return HL_ERR_WORKSPACE;
  lambda_max_kernel<<<n_graphs, kLzThreads, 0, as_stream(stream)>>>(node_ptr, max_nodes, inc_rowptr, inc_edge, tail, head,
                                                                    max_steps, lambda_max, last_change,
                                                                    reinterpret_cast<double*>(workspace));
  HL_LAUNCH_CHECK("lambda_max_kernel");
  return HL_OK;
}
