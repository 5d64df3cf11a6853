// COO -> CSR bucketing (hl_csr_from_coo): one radix sort plus a scatter pass that also derives rowptr from the sorted
// keys.  Runs once per mini-batch and operator; the polynomial SpMM / segment kernels then never touch the int64 COO
// again.  HL_TIE_POSITION: the radix sort is STABLE, so sorting 32-bit keys that hold the row id alone (ceil(log2(rows))
// bits: two 8-bit passes for a 24k-row ZINC batch) leaves every row's entries in COO position order -- a third of the
// passes of the (row, position) 64-bit sort it replaces.  HL_TIE_COLUMN keeps 64-bit (row, column) keys.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace hl {

__global__ void make_keys_kernel(const int64_t* __restrict__ row, const float* __restrict__ row_f32,
                                 const int64_t* __restrict__ col, int64_t nnz, int64_t nrows, int tie,
                                 uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r;
    if (row) {
      r = row[i];
    } else {
      const float f = row_f32[i];
      r = isfinite(f) ? (int64_t)f : -1;                     // +inf cluster id = dropped row
    }
    if (r < 0 || r >= nrows) r = nrows;
    const uint32_t lo = (tie == HL_TIE_COLUMN && col) ? (uint32_t)col[i] : (uint32_t)i;
    keys[i] = ((uint64_t)r << 32) | lo;
    idx[i] = (int32_t)i;
  }
}

__global__ void make_keys32_kernel(const int64_t* __restrict__ row, const float* __restrict__ row_f32, int64_t nnz, int64_t nrows,
                                   uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r;
    if (row) {
      r = row[i];
    } else {
      const float f = row_f32[i];
      r = isfinite(f) ? (int64_t)f : -1;
    }
    if (r < 0 || r >= nrows) r = nrows;
    keys[i] = (uint32_t)r;
    idx[i] = (int32_t)i;
  }
}

__global__ void scatter_sorted32_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ idx,
                                        const int64_t* __restrict__ col, const float* __restrict__ val,
                                        int64_t nnz, int64_t nrows, int32_t* __restrict__ rowptr,
                                        int32_t* __restrict__ colidx, float* __restrict__ vals,
                                        int32_t* __restrict__ perm) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t hi = (int64_t)keys[p];
    const int64_t prev = p > 0 ? (int64_t)keys[p - 1] : -1;
    for (int64_t r = prev + 1; r <= hi; ++r) rowptr[r] = (int32_t)p;
    if (p == nnz - 1)
      for (int64_t r = hi + 1; r <= nrows; ++r) rowptr[r] = (int32_t)nnz;
    if (hi < nrows) {
      const int32_t o = idx[p];
      if (colidx) colidx[p] = col ? (int32_t)col[o] : o;
      if (vals) vals[p] = val[o];
      if (perm) perm[p] = o;
    }
  }
}

__global__ void scatter_sorted_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx,
                                      const int64_t* __restrict__ col, const float* __restrict__ val,
                                      int64_t nnz, int64_t nrows, int32_t* __restrict__ rowptr,
                                      int32_t* __restrict__ colidx, float* __restrict__ vals,
                                      int32_t* __restrict__ perm) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t hi = (int64_t)(keys[p] >> 32);
    const int64_t prev = p > 0 ? (int64_t)(keys[p - 1] >> 32) : -1;
    for (int64_t r = prev + 1; r <= hi; ++r) rowptr[r] = (int32_t)p;
    if (p == nnz - 1)
      for (int64_t r = hi + 1; r <= nrows; ++r) rowptr[r] = (int32_t)nnz;
    if (hi < nrows) {
      const int32_t o = idx[p];
      if (colidx) colidx[p] = col ? (int32_t)col[o] : o;
      if (vals) vals[p] = val[o];
      if (perm) perm[p] = o;
    }
  }
}

static int end_bit_for(int64_t nrows) {
  int b = 0;
  while ((1LL << b) <= nrows) ++b;                            // nrows itself is the "dropped" bucket
  return 32 + b;
}

static size_t cub_sort_bytes(int64_t nnz, int64_t nrows) {
  size_t bytes = 0, bytes32 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, nnz, 0, end_bit_for(nrows));
  cub::DeviceRadixSort::SortPairs(nullptr, bytes32, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, nnz, 0, end_bit_for(nrows) - 32);
  return bytes > bytes32 ? bytes : bytes32;
}

}  // namespace hl

extern "C" size_t hl_csr_from_coo_workspace(int64_t nnz, int64_t nrows) {
  if (nnz < 0 || nrows < 0) return 0;
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  return hl::align_up(n * 8, 256) * 2 + hl::align_up(n * 4, 256) * 2 + hl::align_up(hl::cub_sort_bytes(nnz, nrows), 256) + 256;
}

extern "C" int hl_csr_from_coo(const int64_t* row, const float* row_f32, const int64_t* col, const float* val,
                               int64_t nnz, int64_t nrows, int tie,
                               int32_t* rowptr, int32_t* colidx, float* vals, int32_t* perm,
                               void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  using namespace hl;
  if (nnz < 0 || nrows < 0 || nnz > 0x7fffffffLL || nrows > 0x7ffffffeLL || !rowptr) return HL_ERR_INVALID;
  if ((row == nullptr) == (row_f32 == nullptr)) { if (nnz > 0) return HL_ERR_INVALID; }
  if (vals && !val) return HL_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  if (nnz == 0) {
    HL_CUDA_CHECK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (size_t)(nrows + 1), st));
    return HL_OK;
  }
  if (!workspace || workspace_bytes < hl_csr_from_coo_workspace(nnz, nrows)) return HL_ERR_WORKSPACE;
  char* w = reinterpret_cast<char*>(workspace);
  const size_t n = (size_t)nnz;
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(w);  w += align_up(n * 8, 256);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(w); w += align_up(n * 8, 256);
  int32_t* idx_in = reinterpret_cast<int32_t*>(w);     w += align_up(n * 4, 256);
  int32_t* idx_out = reinterpret_cast<int32_t*>(w);    w += align_up(n * 4, 256);
  size_t cub_bytes = cub_sort_bytes(nnz, nrows);
  void* cub_tmp = w;

  const int threads = 256;
  int64_t want = (nnz + threads - 1) / threads;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  if (tie == HL_TIE_POSITION || !col) {                         // stable sort on the row id alone keeps position order
    uint32_t* k_in = reinterpret_cast<uint32_t*>(keys_in);
    uint32_t* k_out = reinterpret_cast<uint32_t*>(keys_out);
    make_keys32_kernel<<<blocks, threads, 0, st>>>(row, row_f32, nnz, nrows, k_in, idx_in);
    HL_LAUNCH_CHECK("make_keys32_kernel");
    HL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, idx_in, idx_out, nnz, 0,
                                                  end_bit_for(nrows) - 32, st));
    scatter_sorted32_kernel<<<blocks, threads, 0, st>>>(k_out, idx_out, col, val, nnz, nrows, rowptr, colidx, vals, perm);
    HL_LAUNCH_CHECK("scatter_sorted32_kernel");
    return HL_OK;
  }
  make_keys_kernel<<<blocks, threads, 0, st>>>(row, row_f32, col, nnz, nrows, tie, keys_in, idx_in);
  HL_LAUNCH_CHECK("make_keys_kernel");
  HL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, idx_in, idx_out, nnz, 0,
                                                end_bit_for(nrows), st));
  scatter_sorted_kernel<<<blocks, threads, 0, st>>>(keys_out, idx_out, col, val, nnz, nrows, rowptr, colidx, vals, perm);
  HL_LAUNCH_CHECK("scatter_sorted_kernel");
  return HL_OK;
}
