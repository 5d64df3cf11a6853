// Eigenvector positional encodings for a whole mini-batch.
//
// The reference computes, per graph and on the CPU, `eigh` of the dense normalised Laplacians and keeps the
// eigenvectors 1 .. k-1 in ascending eigenvalue order (lib/Hodge_Dataset.py:97-112 `eig_pe`, callers :457-458,
// :586-587, :846-847).  Here every diagonal block of the batch's block-diagonal CSR operator is expanded into a small
// dense matrix and diagonalised by ONE CTA with the cyclic Jacobi method in round-robin ("chess tournament") order:
// in every round the m/2 disjoint index pairs are rotated at once -- the rotation angles from the current 2x2
// pivots, then the row update A <- J^T A and the column updates A <- A J, V <- V J, each a bulk-synchronous pass
// whose accesses stay inside matrix rows (coalesced; the matrices live in L1 / L2-resident scratch).  The iteration
// runs in fp64 (like the Lanczos lambda_max of construct.cu) -- the operator entries are fp32, the encodings are
// rounded to fp32 on the way out -- so the result is the exact spectral decomposition to ~1e-13, independent of
// the summation order; Jacobi needs no tridiagonalisation and converges quadratically (7-10 sweeps).
// Eigenvalues come out unsorted on the diagonal; a rank pass orders them ascending (ties by index) and writes
// eigenvectors rank 1 .. k-1 as the encoding, sign-normalised so that the component of largest magnitude is
// positive (LAPACK's signs are arbitrary and the reference re-draws them at random in every `get`, :429-439).
#include "common.cuh"

namespace hl {

constexpr int kEigThreads = 512;

// A <- dense block g of the CSR operator, V <- I
__global__ void __launch_bounds__(256)
eig_fill_kernel(const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ mat_ptr, const int32_t* __restrict__ rowptr,
                const int32_t* __restrict__ colidx, const float* __restrict__ vals, double* __restrict__ A, double* __restrict__ V) {
  const int g = blockIdx.x;
  const int r0 = seg_ptr[g], n = seg_ptr[g + 1] - r0;
  double* a = A + mat_ptr[g];
  double* v = V + mat_ptr[g];
  const int64_t nn = (int64_t)n * n;
  for (int64_t i = threadIdx.x + (int64_t)blockIdx.y * blockDim.x; i < nn; i += (int64_t)blockDim.x * gridDim.y) {
    a[i] = 0.0;
    v[i] = (i / n == i % n) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(256)
eig_scatter_kernel(const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ mat_ptr, const int32_t* __restrict__ rowptr,
                   const int32_t* __restrict__ colidx, const float* __restrict__ vals, double* __restrict__ A) {
  const int g = blockIdx.x;
  const int r0 = seg_ptr[g], n = seg_ptr[g + 1] - r0;
  double* a = A + mat_ptr[g];
  // one warp per row: duplicate (row, col) entries of a COO-derived CSR add up, in CSR order (no atomics)
  const int warp = (threadIdx.x >> 5) + (blockDim.x >> 5) * blockIdx.y, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x >> 5) * gridDim.y;
  for (int r = warp; r < n; r += nwarps) {
    const int p0 = rowptr[r0 + r], p1 = rowptr[r0 + r + 1];
    for (int c = lane; c < n; c += 32) {
      double s = 0.0;
      bool any = false;
      for (int p = p0; p < p1; ++p)
        if (colidx[p] - r0 == c) { s += (double)vals[p]; any = true; }
      if (any) a[(int64_t)r * n + c] = s;
    }
  }
}

// pair i of round r among m (even) players, circle method: player m-1 stays, the others rotate
__device__ __forceinline__ void eig_pair(int i, int r, int m, int& p, int& q) {
  const int mm = m - 1;
  int a, b;
  if (i == 0) { a = mm; b = r % mm; }
  else { a = (r + i) % mm; b = (r - i + mm) % mm; }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

__global__ void __launch_bounds__(kEigThreads)
eig_jacobi_kernel(const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ mat_ptr, double* __restrict__ A,
                  double* __restrict__ V, int max_sweeps, int32_t* __restrict__ sweeps_out) {
  extern __shared__ double eig_sh[];                      // c[m/2], s[m/2]
  __shared__ double red[kEigThreads / 32];
  __shared__ int any_rot;
  __shared__ double thresh;
  const int g = blockIdx.x;
  const int n = seg_ptr[g + 1] - seg_ptr[g];
  if (n < 2) { if (threadIdx.x == 0 && sweeps_out) sweeps_out[g] = 0; return; }
  double* a = A + mat_ptr[g];
  double* v = V + mat_ptr[g];
  const int m = (n + 1) & ~1;
  const int half = m >> 1;
  double* cs = eig_sh;
  double* sn = eig_sh + half;
  const int tid = threadIdx.x, nt = blockDim.x;

  // Frobenius norm -> absolute pivot threshold (relative-to-diagonal tests fail on the Laplacian's zero eigenvalue)
  double loc = 0.0;
  for (int64_t i = tid; i < (int64_t)n * n; i += nt) loc += a[i] * a[i];
  for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
  if ((tid & 31) == 0) red[tid >> 5] = loc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < nt / 32; ++w) t += red[w];
    thresh = 4e-16 * sqrt(t);                              // off-diagonal mass left behind <= n * thresh: ~1e-13 relative
  }
  __syncthreads();
  const double thr = thresh;

  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) any_rot = 0;
    __syncthreads();
    for (int r = 0; r < m - 1; ++r) {
      // rotation angles of this round's disjoint pairs
      for (int i = tid; i < half; i += nt) {
        int p, q;
        eig_pair(i, r, m, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = a[(int64_t)p * n + q];
          if (fabs(apq) > thr) {
            const double app = a[(int64_t)p * n + p], aqq = a[(int64_t)q * n + q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = rsqrt(1.0 + t * t);
            s = t * c;
            any_rot = 1;
          }
        }
        cs[i] = c;
        sn[i] = s;
      }
      __syncthreads();
      // rows: A <- J^T A   (rows p and q of every pair, all columns)
      for (int idx = tid; idx < half * n; idx += nt) {
        const int i = idx / n, j = idx - i * n;
        const double s = sn[i];
        if (s == 0.0) continue;
        int p, q;
        eig_pair(i, r, m, p, q);
        const double c = cs[i];
        const double x = a[(int64_t)p * n + j], y = a[(int64_t)q * n + j];
        a[(int64_t)p * n + j] = c * x - s * y;
        a[(int64_t)q * n + j] = s * x + c * y;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J   (columns p and q of every pair, all rows; a warp stays inside one row)
      for (int idx = tid; idx < half * n; idx += nt) {
        const int row = idx / half, i = idx - row * half;
        const double s = sn[i];
        if (s == 0.0) continue;
        int p, q;
        eig_pair(i, r, m, p, q);
        const double c = cs[i];
        double* ar = a + (int64_t)row * n;
        double* vr = v + (int64_t)row * n;
        const double x = ar[p], y = ar[q];
        ar[p] = c * x - s * y;
        ar[q] = s * x + c * y;
        const double vx = vr[p], vy = vr[q];
        vr[p] = c * vx - s * vy;
        vr[q] = s * vx + c * vy;
      }
      __syncthreads();
    }
    if (!any_rot) break;                                   // a whole sweep without a pivot above the threshold
    __syncthreads();
  }
  if (tid == 0 && sweeps_out) sweeps_out[g] = sweep;
}

// eigenvalues ascending (ties by index) -> evals[row of the batch]; eigenvectors of rank 1 .. k-1 -> pe[row, 0 .. k-2]
__global__ void __launch_bounds__(256)
eig_extract_kernel(const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ mat_ptr, const double* __restrict__ A,
                   const double* __restrict__ V, int32_t k, float* __restrict__ evals, float* __restrict__ pe, int64_t ld_pe,
                   float* __restrict__ vecs_all) {
  extern __shared__ int eig_rank[];                        // rank -> column index [n], then sign per column [n]
  const int g = blockIdx.x;
  const int r0 = seg_ptr[g], n = seg_ptr[g + 1] - r0;
  const double* a = A + mat_ptr[g];
  const double* v = V + mat_ptr[g];
  int* by_rank = eig_rank;
  float* sign = reinterpret_cast<float*>(eig_rank + n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double li = a[(int64_t)i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double lj = a[(int64_t)j * n + j];
      rank += (lj < li || (lj == li && j < i)) ? 1 : 0;
    }
    by_rank[rank] = i;
    evals[r0 + rank] = (float)li;
    // sign convention: the component of largest magnitude (first one on ties) is positive
    double best = 0.0;
    float sg = 1.f;
    for (int r = 0; r < n; ++r) {
      const double x = v[(int64_t)r * n + i];
      if (fabs(x) > best * (1.0 + 1e-9)) { best = fabs(x); sg = x < 0.0 ? -1.f : 1.f; }
    }
    sign[i] = sg;
  }
  __syncthreads();
  const int cols = k - 1;
  for (int idx = threadIdx.x; idx < n * cols; idx += blockDim.x) {
    const int r = idx / cols, c = idx - r * cols;
    float val = 0.f;                                       // graphs with fewer than k nodes: zero padding (:430-431)
    if (c + 1 < n) {
      const int col = by_rank[c + 1];
      val = (float)v[(int64_t)r * n + col] * sign[col];
    }
    pe[(int64_t)(r0 + r) * ld_pe + c] = val;
  }
  if (vecs_all) {                                          // optional: ALL eigenvectors, columns in ascending order, [n, n] per graph
    float* out = vecs_all + mat_ptr[g];
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int r = idx / n, c = idx - r * n;
      const int col = by_rank[c];
      out[idx] = (float)v[(int64_t)r * n + col] * sign[col];
    }
  }
}

}  // namespace hl

extern "C" size_t hl_eig_pe_workspace(int64_t total_matrix_elements) {
  if (total_matrix_elements < 0) return 0;
  return 2 * hl::align_up((size_t)total_matrix_elements * sizeof(double), 256) + 256;
}

extern "C" int hl_eig_pe(const int32_t* seg_ptr, int32_t n_graphs, int32_t max_n, const int64_t* mat_ptr,
                         int64_t total_matrix_elements, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         int32_t k, float* evals, float* pe, int64_t ld_pe, float* vecs_all, int32_t* sweeps,
                         int32_t max_sweeps, void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (n_graphs < 0 || max_n < 0 || k < 1 || total_matrix_elements < 0 || ld_pe < k - 1) return HL_ERR_INVALID;
  if (n_graphs == 0 || max_n == 0) return HL_OK;
  if (!seg_ptr || !mat_ptr || !rowptr || !colidx || !vals || !evals || (k > 1 && !pe)) return HL_ERR_INVALID;
  if (!workspace || workspace_bytes < hl_eig_pe_workspace(total_matrix_elements)) return HL_ERR_WORKSPACE;
  if (max_sweeps < 1) max_sweeps = 30;
  cudaStream_t st = as_stream(stream);
  double* A = reinterpret_cast<double*>(workspace);
  double* V = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + align_up((size_t)total_matrix_elements * sizeof(double), 256));
  int chunks = (int)(((int64_t)max_n * max_n + 256 * 16 - 1) / (256 * 16));
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  dim3 grid(n_graphs, chunks);
  eig_fill_kernel<<<grid, 256, 0, st>>>(seg_ptr, mat_ptr, rowptr, colidx, vals, A, V);
  HL_LAUNCH_CHECK("eig_fill_kernel");
  int rchunks = (max_n + 7) / 8;
  if (rchunks > 32) rchunks = 32;
  eig_scatter_kernel<<<dim3(n_graphs, rchunks), 256, 0, st>>>(seg_ptr, mat_ptr, rowptr, colidx, vals, A);
  HL_LAUNCH_CHECK("eig_scatter_kernel");
  const int m = (max_n + 1) & ~1;
  const size_t sh = (size_t)m * sizeof(double);
  if (sh > 48 * 1024) return HL_ERR_INVALID;              // > 12k rows per graph: not this kernel's regime
  eig_jacobi_kernel<<<n_graphs, kEigThreads, sh, st>>>(seg_ptr, mat_ptr, A, V, max_sweeps, sweeps);
  HL_LAUNCH_CHECK("eig_jacobi_kernel");
  const size_t sh2 = (size_t)max_n * (sizeof(int) + sizeof(float));
  if (sh2 > 48 * 1024) return HL_ERR_INVALID;
  eig_extract_kernel<<<n_graphs, 256, sh2, st>>>(seg_ptr, mat_ptr, A, V, k, evals, pe, ld_pe, vecs_all);
  HL_LAUNCH_CHECK("eig_extract_kernel");
  return HL_OK;
}
