// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32) by 3xTF32 splitting:
//     C[M,N] = A[M,K] * B[N,K]^T (+ bias | + C)        A, B row-major with K contiguous ("K-major")
// Every fp32 operand x is split as x = hi + lo with hi = x & 0xffffe000 (exact in tf32) and lo = x - hi;
// the product is accumulated as  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  in an fp32 TMEM accumulator (the dropped
// lo*lo term is ~2^-22 relative), which keeps the reference's fp32 parity bar (rtol 1e-4) that plain
// 1xTF32 (2^-11) would break.  This is the dense Theta / MLP feature transform of the path
// (lib/Hodge_Cheb_Conv.py:487,497,509 and :277-288).
//
// Structure (one CTA = one 128 x BN output tile, BN = N <= 256):
//   warp 0      TMA producer: A tile (128x32 fp32, 128B swizzle) + pre-split B_hi / B_lo tiles per k-block
//   warps 2-5   converters: write lo = x - trunc_tf32(x) of the landed A tile into a second buffer (an elementwise
//               pass, so it is oblivious to the swizzled layout); the landed fp32 tile itself serves as hi, because
//               kind::tf32 reads only the sign, exponent and top 10 mantissa bits of each 32-bit element -- then
//               fence.proxy.async and signal the MMA warp; after the main loop the same warps run the
//               epilogue (tcgen05.ld 32x32b -> registers -> bias / accumulate -> 128-bit global stores)
//   warp 1      one elected lane issues 3 x 4 tcgen05.mma (M=128, N=BN, K=8) per k-block into TMEM and
//               releases the stage with tcgen05.commit
// B (the weights, <= 256 x 1408) is split once per call by hl_tf32_split into global hi / lo copies.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace hl {

constexpr int kGmBM = 128;
constexpr int kGmBK = 32;                 // floats per k-block = one 128-byte swizzle span
constexpr int kGmThreads = 320;           // warp 0 TMA, warp 1 MMA issue, warps 2-5 and 6-9: two converter groups
constexpr int kGmConvThreads = 128;       // threads per converter group (one thread per tile row / TMEM lane)

__device__ __forceinline__ uint32_t gm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void gm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gm_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gm_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gm_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GD_%=;\n\t"
      "bra GW_%=;\n\t"
      "GD_%=:\n\t}" ::"r"(gm_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void gm_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          gm_smem_u32(dst)),
      "l"(map), "r"(gm_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared memory tile (written through the generic proxy, fenced) -> global memory through the tensor map; `add`: the tile
// is ADDED to what global memory holds (a reduction performed by the memory system: the SM never reads the old values)
__device__ __forceinline__ void gm_tma_store_2d(const void* src, const CUtensorMap* map, int c0, int c1, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(gm_smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(gm_smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void gm_tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void gm_tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void gm_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = tile row, 32-bit column = k index), B from shared memory
__device__ __forceinline__ void gm_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void gm_tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void gm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}

// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms of 1024 B (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t gm_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// MN-major tf32 operands have exactly one legal shared-memory layout on sm_100: SWIZZLE_128B_BASE32B
// (32-byte chunks swizzled inside the 128-byte row, 4-row / 512-byte atoms; TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  32 MN elements (128 B) per k-row; 4-row k-groups every 512 B (SBO);
// 32-element MN blocks every `mn_block_bytes` (LBO) -- what {32 cols, 32 rows} TMA boxes of a row-major
// [K rows, MN cols] array produce when stacked one after another.
__device__ __forceinline__ uint64_t gm_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t mn_block_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((mn_block_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                       // SWIZZLE_128B_BASE32B
  return d;
}

// BatchNorm statistics of the output, from the epilogue (the Linear -> BatchNorm pairs of the path, lib/Hodge_Cheb_Conv.py:
// 277-288, lib/Hodge_ST_Model.py:578-590): in the coalesced phase of the epilogue a lane holds 8 rows x 4 columns of the
// FINAL values of a 32-row x 32-column block; the 4 lanes that share a column group combine over the block's valid rows
// (two passes over registers: mean, then squared deviations -- no cancellation), and one of them writes mean and M2 of
// the block.  bn_stats_final_tiles_kernel merges the blocks with Chan's formula in fp64, in a fixed order.
__device__ __forceinline__ void gm_bn_block_stats(const float4 (&v)[8], int row_base, int sub_row, int nvalid_rows,
                                                   float* __restrict__ part, int64_t block_index, int32_t N, int col, bool col_ok) {
  const int cnt = max(0, min(32, nvalid_rows - row_base));               // valid rows of this 32-row block (warp-uniform)
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 8; ++it)
    if (row_base + it * 4 + sub_row < nvalid_rows) { s[0] += v[it].x; s[1] += v[it].y; s[2] += v[it].z; s[3] += v[it].w; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 8);
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 16);
  }
  const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
  const float m[4] = {s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv};
  float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 8; ++it)
    if (row_base + it * 4 + sub_row < nvalid_rows) {
      const float d0 = v[it].x - m[0], d1 = v[it].y - m[1], d2 = v[it].z - m[2], d3 = v[it].w - m[3];
      q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q[k] += __shfl_xor_sync(0xffffffffu, q[k], 8);
    q[k] += __shfl_xor_sync(0xffffffffu, q[k], 16);
  }
  if (sub_row == 0 && col_ok) {
    float* p = part + block_index * 2 * (int64_t)N + col;
    *reinterpret_cast<float4*>(p) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(p + N) = make_float4(q[0], q[1], q[2], q[3]);
  }
}

struct GemmParams {
  int32_t M, N, K, bn, stages, tmem_cols;
  const float* bias;
  float* C;
  int64_t ldc;
  int32_t accumulate;
  int32_t kb_first;         // forward mode: k-blocks taken from the first A operand (the rest from the second)
  int32_t k_per_split;      // wgrad mode: rows of the contraction handled by one blockIdx.z (multiple of 32)
  int64_t split_stride;     // wgrad mode: elements between partial outputs of consecutive splits
  int32_t tmem_a_col;       // TS variant: first TMEM column of the A_hi | A_lo slots (64 columns per ring stage)
  int32_t conv_groups;      // converter groups taking alternate ring stages: 1 (192 threads) or 2 (320 threads)
  float* colsum_ws;         // wgrad TS: per-split column sums of G, [splits][M] (the bias gradient falls out of the A pass)
  int32_t num_tiles, tiles_n;   // persistent kernel: output tiles in total / along N
  int32_t acc_stride;           // persistent kernel: TMEM columns between the two accumulators
  int32_t x_tiles_each;         // wgrad with TWO X operands (map_bhi, map_blo): column tiles per operand (0 = one operand)
  float* bn_part;               // forward: per 32-row block BatchNorm statistics of the output [ceil(M/32)][2][N] (mean | M2), or NULL
  const int32_t* bn_nvalid;     // device row count the statistics cover (rows beyond are padding), or NULL = M
  int32_t tma_c;                // persistent kernel: the output has a tensor map (map_c): 32 x 32 chunks leave through TMA
};

// MODE 0: C = A[M,K] B[N,K]^T, both K-major, B pre-split (map_b = hi, map_b2 = lo).
// MODE 1: weight gradient C[M,N] = G[R,M]^T X[R,N] over the row range of blockIdx.z: both operands MN-major
//         (contraction over the rows), both split in the kernel; map_a = G, map_b = X, map_b2 unused.
// TS = 1 (MODE 0 only): the converter warps write A_hi / A_lo into TENSOR MEMORY (tcgen05.st) and the MMAs take A from
//         there; a ring stage then holds only the raw A tile and B_hi / B_lo (48 KB instead of 64 KB at 128 columns:
//         one more stage in flight) and the MMAs read half as many shared-memory bytes.
template <int MODE, int TS = 0>
__global__ void __launch_bounds__(kGmThreads)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_a2,
                   const GemmParams P) {
  extern __shared__ __align__(1024) unsigned char gm_smem[];
  hl::pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, stages = P.stages;
  const uint32_t a_bytes = kGmBM * kGmBK * 4, b_bytes = (uint32_t)bn * kGmBK * 4;
  const uint32_t stage_bytes = (TS ? 1 : 2) * a_bytes + 2 * b_bytes;
  const uint32_t b_off = (TS ? 1 : 2) * a_bytes;                   // B_hi inside a stage (B_lo follows)
  unsigned char* base = gm_smem;                                   // 1024-byte aligned by the launch
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)stages * stage_bytes);
  uint64_t* full_bar = bars;                 // [stages]  TMA bytes landed
  uint64_t* conv_bar = bars + stages;        // [stages]  A split done
  uint64_t* empty_bar = bars + 2 * stages;   // [stages]  MMAs of the stage retired
  uint64_t* acc_bar = bars + 3 * stages;     // [1]       accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * stages + 1);

  const int m0 = blockIdx.x * kGmBM;
  // weight gradient of two activations that share g (dW = g^T [X1 | X2]): column tiles 0 .. x_tiles_each-1 read X1, the rest X2
  const int grp = (MODE == 1 && P.x_tiles_each > 0) ? (int)blockIdx.y / P.x_tiles_each : 0;
  const int n0 = ((int)blockIdx.y - grp * (MODE == 1 ? P.x_tiles_each : 0)) * bn;       // column inside the operand
  const int c_off = grp * P.N;                                                          // column of the operand in the output plane
  const int k_begin = MODE == 1 ? (int)blockIdx.z * P.k_per_split : 0;
  const int k_end = MODE == 1 ? min(P.K, k_begin + P.k_per_split) : P.K;
  const int num_kb = k_end > k_begin ? (k_end - k_begin + kGmBK - 1) / kGmBK : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      gm_mbar_init(&full_bar[s], 1);
      gm_mbar_init(&conv_bar[s], kGmConvThreads);
      gm_mbar_init(&empty_bar[s], 1);
    }
    gm_mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {                                                  // TMEM allocation (whole warp)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  hl::pdl_wait();                                                    // everything above overlapped the predecessor's tail (common.cuh)

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      // ring position as running counters (stage, phase parity): an integer division by the runtime stage count per
      // k-block sat on the serial path of each role
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        gm_mbar_wait(&empty_bar[s], ph ^ 1u);
        unsigned char* st = base + (size_t)s * stage_bytes;
        if (MODE == 0) {
          gm_mbar_arrive_expect_tx(&full_bar[s], a_bytes + 2 * b_bytes);
          // [A | A2] W^T: the second operand continues the contraction (B is packed [N, K1pad + K2])
          if (kb < P.kb_first) gm_tma_load_2d(st, &map_a, &full_bar[s], kb * kGmBK, m0);
          else gm_tma_load_2d(st, &map_a2, &full_bar[s], (kb - P.kb_first) * kGmBK, m0);
          gm_tma_load_2d(st + b_off, &map_bhi, &full_bar[s], kb * kGmBK, n0);
          gm_tma_load_2d(st + b_off + b_bytes, &map_blo, &full_bar[s], kb * kGmBK, n0);
        } else {
          // stack of {32 cols, 32 rows} boxes: 4 for the 128 output rows (G columns), bn/32 for the output columns
          gm_mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes);
          const int row = k_begin + kb * kGmBK;
          if (TS) gm_tma_load_2d(st, &map_a, &full_bar[s], m0, row);       // plain [32 k-rows][128 m] tile for the converters
          else
            for (int j = 0; j < kGmBM / 32; ++j) gm_tma_load_2d(st + j * 4096, &map_a, &full_bar[s], m0 + 32 * j, row);
          const CUtensorMap* mx = grp ? &map_blo : &map_bhi;
          for (int j = 0; j < bn / 32; ++j)
            gm_tma_load_2d(st + b_off + j * 4096, mx, &full_bar[s], n0 + 32 * j, row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, K-major both, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24) |
                             (MODE == 1 ? ((TS ? 0u : (1u << 15)) | (1u << 16)) : 0u);   // A from TMEM is always K-major
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        gm_mbar_wait(&conv_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = gm_smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + a_bytes, b_hi = st + b_off, b_lo = st + b_off + b_bytes;
        if (TS) {
          const uint32_t ta_hi = tmem_base + (uint32_t)P.tmem_a_col + (uint32_t)s * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
          for (int k = 0; k < kGmBK / 8; ++k) {
            const uint64_t dbh = MODE == 0 ? gm_desc_kmajor_sw128(b_hi + (uint32_t)k * 32u) : gm_desc_mnmajor_sw128(b_hi + (uint32_t)k * 1024u, 4096);
            const uint64_t dbl = MODE == 0 ? gm_desc_kmajor_sw128(b_lo + (uint32_t)k * 32u) : gm_desc_mnmajor_sw128(b_lo + (uint32_t)k * 1024u, 4096);
            gm_mma_tf32_ts(tmem_base, ta_lo + 8u * k, dbh, idesc, (kb | k) ? 1u : 0u);   // small terms first
            gm_mma_tf32_ts(tmem_base, ta_hi + 8u * k, dbl, idesc, 1u);
            gm_mma_tf32_ts(tmem_base, ta_hi + 8u * k, dbh, idesc, 1u);
          }
          gm_commit(&empty_bar[s]);
          continue;
        }
#pragma unroll
        for (int k = 0; k < kGmBK / 8; ++k) {
          uint64_t dah, dal, dbh, dbl;
          if (MODE == 0) {
            const uint32_t off = (uint32_t)k * 32u;                 // 8 tf32 = 32 bytes along K inside the swizzle span
            dah = gm_desc_kmajor_sw128(a_hi + off); dal = gm_desc_kmajor_sw128(a_lo + off);
            dbh = gm_desc_kmajor_sw128(b_hi + off); dbl = gm_desc_kmajor_sw128(b_lo + off);
          } else {
            const uint32_t off = (uint32_t)k * 1024u;               // 8 k-rows = one 1 KB atom group
            dah = gm_desc_mnmajor_sw128(a_hi + off, 4096); dal = gm_desc_mnmajor_sw128(a_lo + off, 4096);
            dbh = gm_desc_mnmajor_sw128(b_hi + off, 4096); dbl = gm_desc_mnmajor_sw128(b_lo + off, 4096);
          }
          gm_mma_tf32(tmem_base, dal, dbh, idesc, (kb | k) ? 1u : 0u);   // small terms first
          gm_mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          gm_mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        gm_commit(&empty_bar[s]);                                   // frees the stage when these MMAs retire
      }
      gm_commit(acc_bar);
      hl::pdl_trigger_late();                                        // all MMAs issued: only the epilogue is left
    }
  } else {
    // ------------------------------------ converters, then epilogue ------------------------------------
    // Weight gradient (TS): two converter groups take alternate ring stages -- one stage's chain (transposed LDS of G,
    // split of X in shared memory, tcgen05.st, wait::st) is longer than its MMAs, so a single group was the serial
    // bottleneck of the ring (256 x 704: 61.8 -> 51.7 us).  The forward kernel gains nothing from a second group
    // (92.1 vs 93.4 us, small shapes slightly slower) and runs with one; the groups share the epilogue.
    const int groups = P.conv_groups;
    const int cg = (warp - 2) >> 2;                                  // converter group 0 / 1
    const int ct = (threadIdx.x - 64) & (kGmConvThreads - 1);       // 0..127 inside the group
    float csum = 0.f;                                               // wgrad TS: running sum of this thread's G column
    int s = cg;                                                      // groups <= stages, and stages even when groups == 2
    uint32_t ph = 0;
    for (int kb = cg; kb < num_kb; kb += groups) {
      if (kb != cg) {
        s += groups;
        if (s >= stages) { s -= stages; ph ^= 1u; }
      }
      gm_mbar_wait(&full_bar[s], ph);
      if (TS) {
        // thread = tile row = its own TMEM lane: un-swizzle the 128-byte row (16-byte chunk c sits at c ^ (row % 8)),
        // split, and store hi | lo as 2 x 32 columns of the stage's TMEM slot
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        uint32_t hi[32], lo[32];
        if (MODE == 0) {
          const unsigned char* arow = base + (size_t)s * stage_bytes + (size_t)r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t h = __float_as_uint(xs[t]) & 0xffffe000u;
              hi[4 * c + t] = h;
              lo[4 * c + t] = __float_as_uint(xs[t] - __uint_as_float(h));
            }
          }
        } else {
          // weight gradient: A = G^T.  The G tile sits as [32 k-rows][128 m] (unswizzled); this thread's output row m = r
          // is a column of it (consecutive lanes -> consecutive words: conflict-free)
          const float* acol = reinterpret_cast<const float*>(base + (size_t)s * stage_bytes) + r;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = acol[k * kGmBM];
            csum += x;
            const uint32_t h = __float_as_uint(x) & 0xffffe000u;
            hi[k] = h;
            lo[k] = __float_as_uint(x - __uint_as_float(h));
          }
          float4* b = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + b_off);
          float4* blo = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + b_off + b_bytes);
          for (int idx = ct; idx < (int)(b_bytes / 16); idx += kGmConvThreads) {
            const float4 x = b[idx];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
            blo[idx] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B_lo (generic-proxy writes) -> visible to the MMA
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)P.tmem_a_col + (uint32_t)s * 64u;
        gm_tmem_st_x32(taddr, hi);
        gm_tmem_st_x32(taddr + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        gm_mbar_arrive(&conv_bar[s]);
        continue;
      }
      float4* a = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes);
      float4* alo = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + a_bytes);
#pragma unroll
      for (int i = 0; i < (kGmBM * kGmBK / 4) / kGmConvThreads; ++i) {
        const int idx = ct + i * kGmConvThreads;
        const float4 x = a[idx];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
        alo[idx] = l;                                               // hi stays implicit: kind::tf32 ignores the low 13 mantissa bits
      }
      if (MODE == 1) {
        float4* b = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + b_off);
        float4* blo = reinterpret_cast<float4*>(base + (size_t)s * stage_bytes + b_off + b_bytes);
        for (int idx = ct; idx < (int)(b_bytes / 16); idx += kGmConvThreads) {
          const float4 x = b[idx];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
          blo[idx] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
      gm_mbar_arrive(&conv_bar[s]);
    }
    if (MODE == 1 && TS && P.colsum_ws && blockIdx.y == 0) {
      const int m = m0 + (warp & 3) * 32 + lane;
      if (m < P.M) {                                                 // one plane per split and group
        P.colsum_ws[((int64_t)blockIdx.z * 2 + cg) * P.M + m] = csum;
        if (groups == 1) P.colsum_ws[((int64_t)blockIdx.z * 2 + 1) * P.M + m] = 0.f;
      }
    }
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  (= output rows).  Each 32-column chunk goes TMEM -> registers
    // (thread = row) -> a padded per-warp staging tile in shared memory (the ring is idle by now) -> global memory with
    // 8 lanes per row, so a warp instruction moves four complete 128-byte row segments (full sectors) instead of 32
    // half-filled ones; in accumulate mode the old values are fetched the same way and BEFORE the tcgen05.ld wait.
    int bn_rows = P.M;                                                // BatchNorm statistics: rows they cover (load in flight under the wait)
    if (MODE == 0 && P.bn_part && P.bn_nvalid) bn_rows = min(__ldg(P.bn_nvalid), P.M);
    gm_mbar_wait(acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* Cbase = P.C + (MODE == 1 ? (int64_t)blockIdx.z * P.split_stride : 0);
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    const bool vec_ok = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    constexpr int kPitch = 36;                                        // floats per staged row: 144 B, conflict-free for 16-byte accesses
    float* stage_tile = reinterpret_cast<float*>(base) + (size_t)(warp - 2) * 32 * kPitch;
    const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;          // coalesced phase: 4 rows x 8 float4 per instruction
    for (int c = 32 * cg; c < bn; c += 32 * groups) {                // the groups take alternate 32-column chunks
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      if (vec_ok) {
        const int col = n0 + c + sub_col;
        const bool col_ok = col + 3 < P.N && c + sub_col < bn;
        float4 old[8];
        if (P.accumulate) {                                           // independent loads, in flight under the TMEM read
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = m0 + quad * 32 + it * 4 + sub_row;
            old[it] = (col_ok && rr < P.M) ? *reinterpret_cast<const float4*>(Cbase + (int64_t)rr * P.ldc + c_off + col)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.bias && col_ok) {
          if ((reinterpret_cast<uintptr_t>(P.bias) & 15) == 0) bv = __ldg(reinterpret_cast<const float4*>(P.bias + col));
          else bv = make_float4(__ldg(P.bias + col), __ldg(P.bias + col + 1), __ldg(P.bias + col + 2), __ldg(P.bias + col + 3));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* mine = stage_tile + lane * kPitch;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(mine + j) =
              num_kb > 0 ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        float4 fin[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int lr = it * 4 + sub_row;
          const int rr = m0 + quad * 32 + lr;
          float4 v = *reinterpret_cast<const float4*>(stage_tile + lr * kPitch + sub_col);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if (P.accumulate) { v.x += old[it].x; v.y += old[it].y; v.z += old[it].z; v.w += old[it].w; }
          if (col_ok && rr < P.M) *reinterpret_cast<float4*>(Cbase + (int64_t)rr * P.ldc + c_off + col) = v;
          fin[it] = v;
        }
        if (MODE == 0 && P.bn_part && m0 + quad * 32 < P.M) {             // (warp-uniform) blocks past the last row do not exist
          gm_bn_block_stats(fin, m0 + quad * 32, sub_row, bn_rows, P.bn_part, (int64_t)(m0 / 32 + quad), P.N, col, col_ok);
        }
        __syncwarp();                                                 // the staging tile is reused by the next chunk
        continue;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < P.M) {                                                // unaligned output: element-wise
        float* crow = Cbase + (int64_t)row * P.ldc + c_off + n0 + c;
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c + j;
          if (col >= P.N || c + j >= bn) break;                    // tile overhang / columns past this CTA's bn
          float v = num_kb > 0 ? __uint_as_float(r[j]) : 0.f;
          if (P.bias) v += __ldg(P.bias + col);
          crow[j] = P.accumulate ? crow[j] + v : v;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent forward / data-gradient kernel (MODE 0 with the A operand in tensor memory).
//
// ncu on the one-tile-per-CTA kernel above (profiles/r2_gemm_k256_one_tile_kernel_full.txt): at K = 256 the tensor pipe is active 38 % of
// the time and the SMs sit idle for 29 % of the launch -- L2 and DRAM are far from their limits (13 % / 10 %), the time
// goes into what every CTA pays around its 8 k-blocks: barrier init, TMEM allocation, descriptor fetch, pipeline fill,
// and an epilogue that the converter warps run AFTER the main loop.  Here one CTA per SM walks its share of the output
// tiles: the ring (TMA -> A split into tensor memory -> tcgen05.mma) never drains between tiles, the accumulator is
// DOUBLE-BUFFERED in tensor memory (2 x bn columns + 4 x 64 columns of A slots = all 512), and four dedicated epilogue
// warps drain accumulator b (TMEM -> registers -> padded staging tile -> full 128-byte row segments) while the MMA warp
// already fills accumulator 1 - b with the next tile.
//   warp 0      TMA producer          warp 1      tcgen05.mma issue (one lane)
//   warps 2-5   A converters          warps 6-13  epilogue (two per TMEM lane quadrant, alternate 32-column chunks)
//
// Epilogue of a 32 x 32 chunk: TMEM -> registers (thread = row) -> (+ bias) -> a 4 KB staging tile in the 128-byte
// swizzle -> ONE TMA store of the tile (cp.async.bulk.tensor, clipped at the matrix edge by the tensor map); in
// accumulate mode a TMA reduce-add, so the old values are never read by the SM.  (The register epilogue that came
// before it -- read the tile back 8 lanes per row, fetch the old values, add, store -- kept 16 KB of loads in flight per
// SM and reached 45 % of the HBM floor on the output-bound shapes [229944, 128] x [128, 416..672] of the TSP step,
// tools/dense_shapes_probe.py.)  The register path remains for launches that want the BatchNorm statistics of the final
// values, for the chunk that straddles the end of a column tile, and for outputs a tensor map cannot describe.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPsEpiWarps = 8;
constexpr int kPsThreads = 64 + kGmConvThreads + 32 * kPsEpiWarps;
constexpr int kPsStageFloats = 32 * 32;                       // one staged chunk: 32 rows x 128 B, 128-byte swizzle

__global__ void __launch_bounds__(kPsThreads, 1)
gemm_tf32x3_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                              const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_a2,
                              const __grid_constant__ CUtensorMap map_c, const GemmParams P) {
  extern __shared__ __align__(1024) unsigned char gm_smem[];
  hl::pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, stages = P.stages;
  const uint32_t a_bytes = kGmBM * kGmBK * 4, b_bytes = (uint32_t)bn * kGmBK * 4;
  const uint32_t stage_bytes = a_bytes + 2 * b_bytes;
  unsigned char* base = gm_smem;
  float* staging = reinterpret_cast<float*>(base + (size_t)stages * stage_bytes);          // [8 warps][32 rows][32], 1 KB aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(staging) + kPsEpiWarps * kPsStageFloats * 4);
  uint64_t* full_bar = bars;                 // [stages]  TMA bytes landed
  uint64_t* conv_bar = bars + stages;        // [stages]  A split written to tensor memory
  uint64_t* empty_bar = bars + 2 * stages;   // [stages]  MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * stages;    // [2]       accumulator b complete
  uint64_t* acc_empty = acc_full + 2;        // [2]       accumulator b drained by the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int num_kb = (P.K + kGmBK - 1) / kGmBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      gm_mbar_init(&full_bar[s], 1);
      gm_mbar_init(&conv_bar[s], kGmConvThreads);
      gm_mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      gm_mbar_init(&acc_full[b], 1);
      gm_mbar_init(&acc_empty[b], kPsEpiWarps);                     // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  hl::pdl_wait();                                                    // everything above overlapped the predecessor's tail (common.cuh)

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.tiles_n) * kGmBM, n0 = (tile % P.tiles_n) * bn;
        for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          gm_mbar_wait(&empty_bar[s], ph ^ 1u);
          unsigned char* st = base + (size_t)s * stage_bytes;
          gm_mbar_arrive_expect_tx(&full_bar[s], a_bytes + 2 * b_bytes);
          if (kb < P.kb_first) gm_tma_load_2d(st, &map_a, &full_bar[s], kb * kGmBK, m0);
          else gm_tma_load_2d(st, &map_a2, &full_bar[s], (kb - P.kb_first) * kGmBK, m0);
          gm_tma_load_2d(st + a_bytes, &map_bhi, &full_bar[s], kb * kGmBK, n0);
          gm_tma_load_2d(st + a_bytes + b_bytes, &map_blo, &full_bar[s], kb * kGmBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        gm_mbar_wait(&acc_empty[b], (uint32_t)((it >> 1) & 1) ^ 1u);       // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * P.acc_stride);
        for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          gm_mbar_wait(&conv_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = gm_smem_u32(base + (size_t)s * stage_bytes);
          const uint32_t b_hi = st + a_bytes, b_lo = st + a_bytes + b_bytes;
          const uint32_t ta_hi = tmem_base + (uint32_t)P.tmem_a_col + (uint32_t)s * 64u, ta_lo = ta_hi + 32u;
#pragma unroll
          for (int k = 0; k < kGmBK / 8; ++k) {
            const uint64_t dbh = gm_desc_kmajor_sw128(b_hi + (uint32_t)k * 32u);
            const uint64_t dbl = gm_desc_kmajor_sw128(b_lo + (uint32_t)k * 32u);
            gm_mma_tf32_ts(d_tmem, ta_lo + 8u * k, dbh, idesc, (kb | k) ? 1u : 0u);   // small terms first
            gm_mma_tf32_ts(d_tmem, ta_hi + 8u * k, dbl, idesc, 1u);
            gm_mma_tf32_ts(d_tmem, ta_hi + 8u * k, dbh, idesc, 1u);
          }
          gm_commit(&empty_bar[s]);
        }
        gm_commit(&acc_full[b]);
      }
      hl::pdl_trigger_late();                                        // this CTA's last tile is issued: only its epilogue is left
    }
  } else if (warp < 6) {
    // ------------------------------------ A converters ------------------------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == stages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
        gm_mbar_wait(&full_bar[s], ph);
        uint32_t hi[32], lo[32];
        const unsigned char* arow = base + (size_t)s * stage_bytes + (size_t)r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t h = __float_as_uint(xs[t]) & 0xffffe000u;
            hi[4 * c + t] = h;
            lo[4 * c + t] = __float_as_uint(xs[t] - __uint_as_float(h));
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)P.tmem_a_col + (uint32_t)s * 64u;
        gm_tmem_st_x32(taddr, hi);
        gm_tmem_st_x32(taddr + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        gm_mbar_arrive(&conv_bar[s]);
      }
    }
  } else {
    // ------------------------------------ epilogue ------------------------------------
    const int quad = warp & 3;                                       // TMEM lane quadrant this warp may read
    const int half = (warp - 6) >> 2;                                // the two warps of a quadrant take alternate chunks
    float* stage_tile = staging + (size_t)(warp - 6) * kPsStageFloats;
    unsigned char* stage_row = reinterpret_cast<unsigned char*>(stage_tile) + lane * 128;   // thread = row of the chunk
    const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
    const bool vec_ok = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    const bool bias_vec = P.bias && (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0;
    const int bn_rows = (P.bn_part && P.bn_nvalid) ? min(__ldg(P.bn_nvalid), P.M) : P.M;   // rows the BatchNorm statistics cover
    bool tma_pending = false;                                        // a TMA store may still be reading the staging tile
    int it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const int m0 = (tile / P.tiles_n) * kGmBM, n0 = (tile % P.tiles_n) * bn;
      gm_mbar_wait(&acc_full[b], (uint32_t)((it >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 32 * half; c < bn; c += 64) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * P.acc_stride + c);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        if (vec_ok) {
          // whole chunk inside this column tile (or the tile is the last one: the tensor map clips at column N)
          const bool by_tma = P.tma_c && !P.bn_part && (c + 32 <= bn || n0 + bn >= P.N);
          if (tma_pending) {                                          // the previous store has read the staging tile
            if (lane == 0) gm_tma_store_wait_read();
            __syncwarp();
            tma_pending = false;
          }
          if (by_tma) {
            // bias of column n0 + c + lane, fetched under the TMEM read and handed round by shuffles (thread = row needs all 32)
            float bias_l = 0.f;
            if (P.bias && n0 + c + lane < P.N) bias_l = __ldg(P.bias + n0 + c + lane);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                     __uint_as_float(r[4 * j + 3]));
              if (P.bias) {
                v.x += __shfl_sync(0xffffffffu, bias_l, 4 * j);
                v.y += __shfl_sync(0xffffffffu, bias_l, 4 * j + 1);
                v.z += __shfl_sync(0xffffffffu, bias_l, 4 * j + 2);
                v.w += __shfl_sync(0xffffffffu, bias_l, 4 * j + 3);
              }
              *reinterpret_cast<float4*>(stage_row + ((j ^ (lane & 7)) << 4)) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0 && m0 + quad * 32 < P.M && n0 + c < P.N) gm_tma_store_2d(stage_tile, &map_c, n0 + c, m0 + quad * 32, P.accumulate != 0);
            tma_pending = true;
            continue;
          }
          const int col = n0 + c + sub_col;
          const bool col_ok = col + 3 < P.N && c + sub_col < bn;
          float4 old[8];
          if (P.accumulate) {
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int rr = m0 + quad * 32 + i8 * 4 + sub_row;
              old[i8] = (col_ok && rr < P.M) ? *reinterpret_cast<const float4*>(P.C + (int64_t)rr * P.ldc + col)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias && col_ok) {
            if (bias_vec) bv = __ldg(reinterpret_cast<const float4*>(P.bias + col));
            else bv = make_float4(__ldg(P.bias + col), __ldg(P.bias + col + 1), __ldg(P.bias + col + 2), __ldg(P.bias + col + 3));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stage_row + ((j ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          float4 fin[8];
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int lr = i8 * 4 + sub_row;
            const int rr = m0 + quad * 32 + lr;
            float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(stage_tile) + lr * 128 +
                                                        (((lane & 7) ^ (lr & 7)) << 4));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (P.accumulate) { v.x += old[i8].x; v.y += old[i8].y; v.z += old[i8].z; v.w += old[i8].w; }
            if (col_ok && rr < P.M) *reinterpret_cast<float4*>(P.C + (int64_t)rr * P.ldc + col) = v;
            fin[i8] = v;
          }
          if (P.bn_part && m0 + quad * 32 < P.M) {
            gm_bn_block_stats(fin, m0 + quad * 32, sub_row, bn_rows, P.bn_part, (int64_t)(m0 / 32 + quad), P.N, col, col_ok);
          }
          __syncwarp();
          continue;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = m0 + quad * 32 + lane;
        if (row < P.M) {
          float* crow = P.C + (int64_t)row * P.ldc + n0 + c;
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + c + j;
            if (col >= P.N || c + j >= bn) break;
            float v = __uint_as_float(r[j]);
            if (P.bias) v += __ldg(P.bias + col);
            crow[j] = P.accumulate ? crow[j] + v : v;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) gm_mbar_arrive(&acc_empty[b]);                  // this warp's chunks of accumulator b are free again
    }
    if (tma_pending && lane == 0) gm_tma_store_wait_all();           // the staging tile must outlive the last store
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

// x -> hi (low 13 mantissa bits cleared) and lo = x - hi; optional transpose: out[c, r] = split(src[r, c])
__global__ void tf32_split_kernel(const float* __restrict__ src, int64_t ld_src, int32_t rows, int32_t cols, int transpose,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ld_out) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float x = src[(int64_t)r * ld_src + c];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int64_t o = transpose ? (int64_t)c * ld_out + r : (int64_t)r * ld_out + c;
    hi[o] = h;
    lo[o] = x - h;
  }
}

// the same split for a whole table of weight views in one launch: blockIdx.y = table entry
__global__ void __launch_bounds__(256)
tf32_split_batch_kernel(const hl_split_desc* __restrict__ table) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const hl_split_desc D = table[blockIdx.y];
  const int64_t n = (int64_t)D.rows * D.cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D.cols), c = (int)(i - (int64_t)r * D.cols);
    const float x = D.src[(int64_t)r * D.ld_src + c];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int64_t o = D.transpose ? (int64_t)c * D.ld_out + r : (int64_t)r * D.ld_out + c;
    D.hi[o] = h;
    D.lo[o] = x - h;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements); box = {32 cols, box_rows}; 128-byte swizzle
static bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                     int box_cols = kGmBK, bool mn_major = false, bool no_swizzle = false) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : (mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace hl

extern "C" int hl_tf32_split(const float* src, int64_t ld_src, int32_t rows, int32_t cols, int transpose,
                             float* hi, float* lo, int64_t ld_out, hl_stream_t stream) {
  using namespace hl;
  if (rows < 0 || cols < 0 || !hi || !lo) return HL_ERR_INVALID;
  if (rows == 0 || cols == 0) return HL_OK;
  if (!src) return HL_ERR_INVALID;
  const int64_t n = (int64_t)rows * cols;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  hl::launch_pdl(tf32_split_kernel, (int)blocks, 256, 0, as_stream(stream), src, ld_src, rows, cols, transpose, hi, lo, ld_out);
  HL_LAUNCH_CHECK("tf32_split_kernel");
  return HL_OK;
}

extern "C" int hl_tf32_split_batch(const hl_split_desc* table, int32_t n_entries, int64_t max_elements, hl_stream_t stream) {
  using namespace hl;
  if (n_entries < 0 || max_elements < 0) return HL_ERR_INVALID;
  if (n_entries == 0 || max_elements == 0) return HL_OK;
  if (!table || n_entries > 65535) return HL_ERR_INVALID;
  int64_t bx = (max_elements + 255) / 256;
  if (bx > 32) bx = 32;
  hl::launch_pdl(tf32_split_batch_kernel, dim3((unsigned)bx, (unsigned)n_entries), 256, 0, as_stream(stream), table);
  HL_LAUNCH_CHECK("tf32_split_batch_kernel");
  return HL_OK;
}

extern "C" int hl_gemm2_bn_tf32x3(const float* A, int64_t lda, int32_t K, const float* A2, int64_t lda2, int32_t K2,
                                  const float* Bhi, const float* Blo, int64_t ldb, int32_t M, int32_t N,
                                  const float* bias, float* C, int64_t ldc, int accumulate, float* bn_part,
                                  const int32_t* bn_nvalid, hl_stream_t stream);

// returns HL_OK, or 1 when the shape / alignment is not supported (caller uses a library GEMM instead)
extern "C" int hl_gemm_tf32x3(const float* A, int64_t lda, const float* Bhi, const float* Blo, int64_t ldb,
                              int32_t M, int32_t N, int32_t K, const float* bias, float* C, int64_t ldc,
                              int accumulate, hl_stream_t stream) {
  return hl_gemm2_bn_tf32x3(A, lda, K, nullptr, 0, 0, Bhi, Blo, ldb, M, N, bias, C, ldc, accumulate, nullptr, nullptr, stream);
}

extern "C" int hl_gemm2_tf32x3(const float* A, int64_t lda, int32_t K, const float* A2, int64_t lda2, int32_t K2,
                               const float* Bhi, const float* Blo, int64_t ldb, int32_t M, int32_t N,
                               const float* bias, float* C, int64_t ldc, int accumulate, hl_stream_t stream) {
  return hl_gemm2_bn_tf32x3(A, lda, K, A2, lda2, K2, Bhi, Blo, ldb, M, N, bias, C, ldc, accumulate, nullptr, nullptr, stream);
}

extern "C" size_t hl_gemm_bn_part_floats(int32_t M, int32_t N) {
  if (M < 0 || N < 0) return 0;
  return (size_t)((M + 31) / 32) * 2 * (size_t)N;
}

// C = [A1 | A2] * B^T: B = [N, pad32(K1) + K2] (the columns of the second block start at pad32(K1)).
// bn_part (nullable, [ceil(M/32)][2][N] floats): BatchNorm statistics of the FINAL output values per 32-row block (mean | M2
// over the rows below *bn_nvalid), written by the epilogue -- pass it on the last launch that touches C; needs 16-byte aligned
// C rows (returns 1 otherwise, like every unsupported shape).
extern "C" int hl_gemm2_bn_tf32x3(const float* A, int64_t lda, int32_t K, const float* A2, int64_t lda2, int32_t K2,
                                  const float* Bhi, const float* Blo, int64_t ldb, int32_t M, int32_t N,
                                  const float* bias, float* C, int64_t ldc, int accumulate, float* bn_part,
                                  const int32_t* bn_nvalid, hl_stream_t stream) {
  using namespace hl;
  if (M < 0 || N < 1 || K < 1 || K2 < 0 || !C) return HL_ERR_INVALID;
  if (M == 0) return HL_OK;
  if (!A || !Bhi || !Blo || (K2 > 0 && !A2)) return HL_ERR_INVALID;
  if (K2 > 0 && (lda2 % 4 != 0 || !aligned_to(A2, 16))) return 1;
  const int kb_first = (K + kGmBK - 1) / kGmBK;
  const int Ktot = K2 > 0 ? kb_first * kGmBK + K2 : K;
  // TMA: 16-byte aligned base and row pitch; MMA: N multiple of 16, one N tile of <= 256 columns per CTA
  if (lda % 4 != 0 || ldb % 4 != 0 || !aligned_to(A, 16) || !aligned_to(Bhi, 16) || !aligned_to(Blo, 16)) return 1;
  if (N % 16 != 0) return 1;
  if (bn_part && (ldc % 4 != 0 || !aligned_to(C, 16) || !aligned_to(bn_part, 16))) return 1;
  // Column tile: as wide as possible (<= 256 columns, multiple of 16; TMA zero-fills the overhang).  Narrower
  // tiles with two co-resident CTAs per SM were measured slower (A re-read from L2, N=64 MMAs): 181 vs 155 us
  // on [24000,1408]x[1408,256].
  // Column tile: <= 128 columns (3 ring stages of 64 KB) measured best: 141 us vs 155 us (256 columns, 2 stages)
  // vs 217 us (64 columns) on [24000,1408]x[1408,256].  An L2 prefetch of the activations 8 k-blocks ahead made
  // no difference (the ring is not DRAM-latency bound).  HL_GEMM_MAX_BN overrides for experiments.
  static int max_bn = 0;
  if (max_bn == 0) {
    const char* e = getenv("HL_GEMM_MAX_BN");
    max_bn = e ? atoi(e) : 128;
    if (max_bn < 16 || max_bn > 256) max_bn = 128;
  }
  const int ntiles = (N + max_bn - 1) / max_bn;
  const int bn = ((N + ntiles - 1) / ntiles + 15) / 16 * 16;
  // TS variant (A_hi / A_lo in tensor memory): HL_GEMM_TS=0 selects the all-shared-memory kernel
  static int use_ts = -1;
  if (use_ts < 0) { const char* e = getenv("HL_GEMM_TS"); use_ts = e ? atoi(e) : 1; }
  const size_t stage_bytes = (use_ts ? 1 : 2) * (size_t)kGmBM * kGmBK * 4 + 2 * (size_t)bn * kGmBK * 4;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) return 1;
  // a ring deeper than the k-loop only costs shared memory; and when the grid is between one and two waves
  // (ZINC: 189-205 row tiles on 148 SMs) a ring that fits twice per SM lets all CTAs be resident at once, so the
  // fixed per-CTA cost (TMEM allocation, descriptor fetch, pipeline fill, epilogue) of the "second wave" overlaps
  const int num_kb = (Ktot + kGmBK - 1) / kGmBK;
  if (stages > num_kb) stages = num_kb < 2 ? 2 : num_kb;
  static int co_resident = -1;
  if (co_resident < 0) { const char* e = getenv("HL_GEMM_CORESIDENT"); co_resident = e ? atoi(e) : 1; }
  const int64_t ctas = (int64_t)((M + kGmBM - 1) / kGmBM) * ntiles;
  const int tmem_a_col = (bn + 31) / 32 * 32;
  if (co_resident && ctas > 148 && (ctas <= 2 * 148 || co_resident == 2)) {
    int s2 = (int)((108 * 1024) / stage_bytes);
    if (use_ts)                                              // two resident CTAs share the 512 TMEM columns
      while (s2 >= 2 && tmem_a_col + 64 * s2 > 256) --s2;
    if (s2 >= 2 && stages > s2) stages = s2;
  }
  if (use_ts)
    while (stages > 2 && tmem_a_col + 64 * stages > 512) --stages;
  int tmem_cols = 32;
  while (tmem_cols < (use_ts ? tmem_a_col + 64 * stages : bn)) tmem_cols <<= 1;
  const size_t smem = stages * stage_bytes + (3 * stages + 2) * sizeof(uint64_t) + 1024;

  CUtensorMap ma, mbh, mbl, ma2;
  if (!make_map(&ma, A, M, K, lda, kGmBM) || !make_map(&mbh, Bhi, N, Ktot, ldb, bn) || !make_map(&mbl, Blo, N, Ktot, ldb, bn))
    return 1;
  if (K2 > 0) { if (!make_map(&ma2, A2, M, K2, lda2, kGmBM)) return 1; }
  else ma2 = ma;
  static DeviceOnce configured;                                     // the attribute is per device
  if (configured.need()) {
    HL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  // persistent variant (default; HL_GEMM_PERSISTENT=0 selects the one-tile-per-CTA kernel): needs the A operand in tensor
  // memory and two accumulators + the A slots inside the 512 TMEM columns
  static int persistent = -1;
  if (persistent < 0) { const char* e = getenv("HL_GEMM_PERSISTENT"); persistent = e ? atoi(e) : 1; }
  // Every forward / data-gradient launch takes it.  Before the TMA-store epilogue the one-tile kernel was faster up to the
  // 296 tiles it keeps resident at once ([24144,64] x [64,64]: 5.6 vs 6.4 us) and HL_GEMM_PERSISTENT=2 restores that rule;
  // with eight epilogue warps and asynchronous stores the persistent kernel wins there too -- same-box A/B of the whole
  // step, 4 alternating runs: ZINC 4.446 -> 4.382 ms, CIFAR 15.22 -> 15.04, zinc_default 10.90 -> 10.79, TSP unchanged.
  if (persistent && use_ts && bn <= 128 && (persistent != 2 || ctas > 2 * 148)) {
    const int acc_stride = (bn + 31) / 32 * 32;
    int ps = (int)((227 * 1024 - 1024 - kPsEpiWarps * kPsStageFloats * 4 - 256) / stage_bytes);
    if (ps > 6) ps = 6;
    while (ps > 2 && 2 * acc_stride + 64 * ps > 512) --ps;
    if (ps > num_kb && num_kb >= 2) ps = num_kb;
    if (ps >= 2 && 2 * acc_stride + 64 * ps <= 512) {
      int cols = 32;
      while (cols < 2 * acc_stride + 64 * ps) cols <<= 1;
      const size_t smem_ps = (size_t)ps * stage_bytes + kPsEpiWarps * kPsStageFloats * 4 + (3 * ps + 5) * sizeof(uint64_t) + 1024;
      CUtensorMap pa, pbh, pbl, pa2, pc;
      // the output through a tensor map of 32 x 32 boxes (HL_GEMM_TMA_STORE=0: the register epilogue for every chunk)
      static int tma_store = -1;
      if (tma_store < 0) { const char* e = getenv("HL_GEMM_TMA_STORE"); tma_store = e ? atoi(e) : 1; }
      const bool c_map = tma_store && ldc % 4 == 0 && aligned_to(C, 16) && make_map(&pc, C, M, N, ldc, 32, 32);
      if (!make_map(&pa, A, M, K, lda, kGmBM) || !make_map(&pbh, Bhi, N, Ktot, ldb, bn) || !make_map(&pbl, Blo, N, Ktot, ldb, bn))
        return 1;
      if (K2 > 0) { if (!make_map(&pa2, A2, M, K2, lda2, kGmBM)) return 1; }
      else pa2 = pa;
      static DeviceOnce configured_ps;
      if (configured_ps.need())
        HL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      GemmParams Q;
      Q.M = M; Q.N = N; Q.K = Ktot; Q.bn = bn; Q.stages = ps; Q.tmem_cols = cols;
      Q.bias = bias; Q.C = C; Q.ldc = ldc; Q.accumulate = accumulate; Q.k_per_split = 0; Q.split_stride = 0;
      Q.tmem_a_col = 2 * acc_stride; Q.colsum_ws = nullptr; Q.conv_groups = 1;
      Q.kb_first = K2 > 0 ? kb_first : 0x7fffffff;
      Q.x_tiles_each = 0; Q.bn_part = bn_part; Q.bn_nvalid = bn_nvalid; Q.tma_c = c_map ? 1 : 0;
      Q.tiles_n = ntiles; Q.num_tiles = ((M + kGmBM - 1) / kGmBM) * ntiles; Q.acc_stride = acc_stride;
      const int sms = device_sm_count();
      const int grid_ps = Q.num_tiles < sms ? Q.num_tiles : sms;
      if (!c_map) pc = pa;
      hl::launch_pdl(gemm_tf32x3_persistent_kernel, grid_ps, kPsThreads, smem_ps, as_stream(stream), pa, pbh, pbl, pa2, pc, Q);
      HL_LAUNCH_CHECK("gemm_tf32x3_persistent_kernel");
      return HL_OK;
    }
  }
  GemmParams P;
  P.M = M; P.N = N; P.K = Ktot; P.bn = bn; P.stages = stages; P.tmem_cols = tmem_cols;
  P.num_tiles = 0; P.tiles_n = 1; P.acc_stride = 0; P.x_tiles_each = 0; P.bn_part = bn_part; P.bn_nvalid = bn_nvalid; P.tma_c = 0;
  P.bias = bias; P.C = C; P.ldc = ldc; P.accumulate = accumulate; P.k_per_split = 0; P.split_stride = 0;
  P.tmem_a_col = tmem_a_col; P.colsum_ws = nullptr; P.conv_groups = 1;
  P.kb_first = K2 > 0 ? kb_first : 0x7fffffff;
  dim3 grid((M + kGmBM - 1) / kGmBM, ntiles);
  if (use_ts) hl::launch_pdl(gemm_tf32x3_kernel<0, 1>, grid, 64 + kGmConvThreads, smem, as_stream(stream), ma, mbh, mbl, ma2, P);
  else hl::launch_pdl(gemm_tf32x3_kernel<0, 0>, grid, 64 + kGmConvThreads, smem, as_stream(stream), ma, mbh, mbl, ma2, P);
  HL_LAUNCH_CHECK("gemm_tf32x3_kernel");
  return HL_OK;
}

// dw (=|+=) sum over the row splits: 8 lanes per group of 4 consecutive output elements (one float4 per split
// plane and lane: k = lane, lane + 8, ...; whole 32-byte sectors of every plane), fixed shuffle tree -> deterministic.
// One call handles output group i (all 8 lanes of the group call it together).
__device__ __forceinline__ void gm_split_reduce_group(const float* __restrict__ partial, int32_t splits, int64_t split_stride,
                                                      int32_t fo, int32_t fi, float* __restrict__ dw, int64_t ld_dw, int accumulate,
                                                      const float* __restrict__ cs_partial, float* __restrict__ dbias,
                                                      int accumulate_bias, float* __restrict__ dw2, int64_t ld_dw2,
                                                      int32_t fi_first, int64_t i) {
  // fi: columns of a partial plane; with dw2 the plane is [dW1 (fi_first columns) | dW2] and goes to two destinations
  const int64_t n4 = ((int64_t)fo * fi) >> 2;                      // fi % 32 == 0: rows never straddle a float4
  const int64_t c4 = dbias ? (fo >> 2) : 0;                         // + the [splits][fo] column sums of G (bias gradient)
  const int64_t tot4 = n4 + c4;
  const int sub = threadIdx.x & 7;
  const bool vec_out = (ld_dw % 4 == 0) && ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) &&
                       (!dw2 || ((ld_dw2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(dw2) & 15) == 0)));
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < tot4) {
    const bool is_cs = i >= n4;
    const float4* p = is_cs ? reinterpret_cast<const float4*>(cs_partial) + (i - n4) : reinterpret_cast<const float4*>(partial) + i;
    const int64_t stride4 = is_cs ? (int64_t)(fo >> 2) : (split_stride >> 2);
    int k = sub;
    const int planes = is_cs ? 2 * splits : splits;               // column sums: one plane per split and converter group
    for (; k + 24 < planes; k += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p + (int64_t)(k + 8 * u) * stride4);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; k < planes; k += 8) {
      const float4 v = __ldg(p + (int64_t)k * stride4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
#pragma unroll
  for (int m = 1; m < 8; m <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, m);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, m);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, m);
    s.w += __shfl_xor_sync(0xffffffffu, s.w, m);
  }
  if (sub == 0 && i >= n4 && i < tot4) {
    float* q = dbias + ((i - n4) << 2);
    q[0] = accumulate_bias ? q[0] + s.x : s.x; q[1] = accumulate_bias ? q[1] + s.y : s.y;
    q[2] = accumulate_bias ? q[2] + s.z : s.z; q[3] = accumulate_bias ? q[3] + s.w : s.w;
  }
  if (sub == 0 && i < n4) {
    const int64_t e = i << 2;
    const int64_t o = e / fi, c = e - o * fi;
    float* q = (dw2 && c >= fi_first) ? dw2 + o * ld_dw2 + (c - fi_first) : dw + o * ld_dw + c;
    if (vec_out) {
      float4* q4 = reinterpret_cast<float4*>(q);
      if (accumulate) { const float4 t = *q4; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
      *q4 = s;
    } else {
      q[0] = accumulate ? q[0] + s.x : s.x; q[1] = accumulate ? q[1] + s.y : s.y;
      q[2] = accumulate ? q[2] + s.z : s.z; q[3] = accumulate ? q[3] + s.w : s.w;
    }
  }
}

__device__ __forceinline__ int64_t gm_split_reduce_groups(int32_t fo, int32_t fi, bool with_bias) {
  const int64_t tot4 = (((int64_t)fo * fi) >> 2) + (with_bias ? (fo >> 2) : 0);
  return (tot4 + 3) & ~(int64_t)3;                                // whole shuffle groups of a warp
}

__global__ void __launch_bounds__(256)
gm_split_reduce_kernel(const float* __restrict__ partial, int32_t splits, int64_t split_stride, int32_t fo,
                       int32_t fi, float* __restrict__ dw, int64_t ld_dw, int accumulate,
                       const float* __restrict__ cs_partial, float* __restrict__ dbias, int accumulate_bias,
                       float* __restrict__ dw2, int64_t ld_dw2, int32_t fi_first) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int64_t groups = gm_split_reduce_groups(fo, fi, dbias != nullptr);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < groups; i += ((int64_t)gridDim.x * blockDim.x) >> 3)
    gm_split_reduce_group(partial, splits, split_stride, fo, fi, dw, ld_dw, accumulate, cs_partial, dbias, accumulate_bias, dw2,
                          ld_dw2, fi_first, i);
}

// The reduces of MANY weight gradients in one launch (hl_wgrad_reduce_batch): the descriptors travel as kernel
// parameters (constant bank, no table upload); work unit = 32 output groups (one 256-thread block pass), the units of
// descriptor d are [block_start[d], block_start[d + 1]).  Same arithmetic and order as gm_split_reduce_kernel.
constexpr int kReduceBatchMax = 192;                               // 192 x 96 B + header < the 32 KB parameter limit
struct ReduceBatch {
  int32_t n, total_blocks;
  hl_wgrad_reduce_desc d[kReduceBatchMax];
};
__global__ void __launch_bounds__(256) gm_split_reduce_batch_kernel(const __grid_constant__ ReduceBatch B) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  for (int blk = blockIdx.x; blk < B.total_blocks; blk += gridDim.x) {
    int lo = 0, hi = B.n - 1;
    while (lo < hi) {                                                // last descriptor whose first unit is <= blk
      const int mid = (lo + hi + 1) >> 1;
      if (B.d[mid].block_start <= blk) lo = mid; else hi = mid - 1;
    }
    const hl_wgrad_reduce_desc& D = B.d[lo];
    const int64_t i = ((int64_t)(blk - D.block_start) * 256 + threadIdx.x) >> 3;
    if (i < gm_split_reduce_groups(D.fo, D.fi, D.dbias != nullptr))   // uniform per group of 8 lanes and per warp of 4 groups
      gm_split_reduce_group(D.partial, D.splits, D.split_stride, D.fo, D.fi, D.dw, D.ld_dw, D.accumulate, D.cs_partial, D.dbias,
                            D.accumulate_bias, D.dw2, D.ld_dw2, D.fi_first, i);
  }
}

// row splits of the weight gradient: tiles x splits CTAs fill `waves` rounds of the 148 SMs without spilling into
// another one (one CTA per SM: the ring takes most of the shared memory).  One round measured best
// (tools/wgrad_shapes_probe.py, sum over the ZINC stack: 0.70 ms vs 0.80 ms for two rounds, 0.89 ms for three): half
// as many partial tiles to write and to reduce, and a k-loop twice as long per prologue / epilogue.
static int wgrad_tc_splits(int32_t nrows, int tiles) {
  static int waves = 0;
  if (waves == 0) { const char* e = getenv("HL_WGRAD_WAVES"); waves = e ? atoi(e) : 1; if (waves < 1 || waves > 4) waves = 1; }
  static int ctas = -1;                                 // CTAs one weight-gradient launch spreads over (tiles x splits)
  if (ctas < 0) { const char* e = getenv("HL_WGRAD_CTAS"); ctas = e ? atoi(e) : 148; if (ctas < 1) ctas = 148; }
  int s = (waves * ctas) / tiles;
  // Weight gradients run on auxiliary streams beside the critical chain of the step; filling all 148 SMs with them takes
  // SMs from that chain and multiplies the partial tiles the reduce has to read.  Measured inside the ZINC step: 148 CTAs
  // per launch 4.77 ms, at most 16 row splits per tile 4.57 ms (8: 5.04 ms -- then the weight gradients themselves become
  // the tail).  Long batches keep more splits so that no CTA walks more than ~4096 rows.  HL_WGRAD_SPLIT_CAP overrides.
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("HL_WGRAD_SPLIT_CAP"); cap = e ? atoi(e) : 16; }
  if (cap > 0) {
    const int by_rows = (nrows + 4095) / 4096;
    const int lim = cap > by_rows ? cap : by_rows;
    if (s > lim) s = lim;
  }
  const int max_s = (nrows + 255) / 256;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

static int wgrad_max_bn() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("HL_WGRAD_MAX_BN");
    v = e ? atoi(e) : 256;
    if (v < 32 || v > 256 || v % 32 != 0) v = 256;
  }
  return v;
}

// column tiles of the weight gradient: 128 columns (64 KB / 48 KB ring stages, 3-4 in flight) from 256 input features
// on, one tile of up to 256 columns below (measured, tools/wgrad_shapes_probe.py: 86.5 -> 75.5 us at 256 x 704,
// 48 -> 37.6 us at 256 x 256, but 24.9 -> 30.2 us at 128 x 192 when split in two 96-column tiles)
static int wgrad_ntiles(int32_t fi) {
  const char* e = getenv("HL_WGRAD_MAX_BN");
  const int max_bn = e ? wgrad_max_bn() : (fi >= 256 ? 128 : 256);
  return (fi + max_bn - 1) / max_bn;
}

extern "C" size_t hl_wgrad_tf32x3_workspace(int32_t nrows, int32_t fo, int32_t fi) {
  if (nrows < 0 || fo < 1 || fi < 1) return 0;
  const int ntiles = wgrad_ntiles(fi);
  const int tiles = ((fo + hl::kGmBM - 1) / hl::kGmBM) * ntiles;
  const size_t splits = (size_t)wgrad_tc_splits(nrows, tiles);
  return hl::align_up(splits * (size_t)fo * (size_t)fi * sizeof(float), 256) + 2 * splits * (size_t)fo * sizeof(float) + 256;
}

extern "C" int hl_wgrad_bias_tf32x3(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo,
                                    int32_t fi, float* dw, int64_t ld_dw, int accumulate, float* dbias, int accumulate_bias,
                                    void* workspace, size_t workspace_bytes, hl_stream_t stream);

// dW[fo,fi] (=|+=) g[R,fo]^T x[R,fi] on the tensor cores (3xTF32), split over row ranges, partial tiles summed
// in a fixed order.  Returns 1 when the shape is unsupported (caller falls back to hl_wgrad).
extern "C" int hl_wgrad_tf32x3(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo,
                               int32_t fi, float* dw, int64_t ld_dw, int accumulate, void* workspace,
                               size_t workspace_bytes, hl_stream_t stream) {
  return hl_wgrad_bias_tf32x3(g, ld_g, x, ld_x, nrows, fo, fi, dw, ld_dw, accumulate, nullptr, 0, workspace, workspace_bytes,
                              stream);
}

// ... and, when dbias != NULL, dbias[fo] (=|+=) column sums of g in the same two launches: the converter warps that move
// G^T into tensor memory add up their column as they go.  Returns 2 (dW done, dbias NOT done: use hl_colsum) when the
// variant in use cannot fold the bias in.
// nx = 2 (x2 != NULL): two activations of the same shape that share g -- dW1 = g^T x1, dW2 = g^T x2 from ONE launch (column
// tiles 0 .. ntiles-1 read x1, the rest x2; the partial planes are [fo][2 fi]) and ONE reduce with two destinations.
static int wgrad_launch(const float* g, int64_t ld_g, const float* x, int64_t ld_x, const float* x2, int64_t ld_x2,
                        int32_t nrows, int32_t fo, int32_t fi, float* dw, int64_t ld_dw, float* dw2, int64_t ld_dw2,
                        int accumulate, float* dbias, int accumulate_bias, void* workspace, size_t workspace_bytes,
                        hl_stream_t stream, hl_wgrad_reduce_desc* defer = nullptr) {
  using namespace hl;
  const int nx = x2 ? 2 : 1;
  if (nrows < 0 || fo < 1 || fi < 1 || !dw || (x2 && !dw2)) return HL_ERR_INVALID;
  if (nrows > 0 && (!g || !x)) return HL_ERR_INVALID;
  if (ld_g % 4 != 0 || ld_x % 4 != 0 || !aligned_to(g, 16) || !aligned_to(x, 16)) return 1;
  if (x2 && (ld_x2 % 4 != 0 || !aligned_to(x2, 16))) return 1;
  if (fo % 4 != 0 || fi % 32 != 0 || nrows < 512) return 1;
  const int ntiles = wgrad_ntiles(fi);
  const int bn = ((fi + ntiles - 1) / ntiles + 31) / 32 * 32;     // multiple of 32: whole {32 x 32} boxes
  const int mtiles = (fo + kGmBM - 1) / kGmBM;
  const int splits = wgrad_tc_splits(nrows, mtiles * ntiles * nx);
  const int64_t fi_tot = (int64_t)fi * nx;
  const size_t need = align_up((size_t)splits * (size_t)fo * (size_t)fi_tot * sizeof(float), 256) + 2 * (size_t)splits * (size_t)fo * sizeof(float) + 256;
  if (!workspace || workspace_bytes < need) return HL_ERR_WORKSPACE;
  static int use_ts = -1;
  if (use_ts < 0) { const char* e = getenv("HL_WGRAD_TS"); use_ts = e ? atoi(e) : 1; }
  if (nx == 2 && !use_ts) return 1;
  const size_t stage_bytes = (use_ts ? 1 : 2) * (size_t)kGmBM * kGmBK * 4 + 2 * (size_t)bn * kGmBK * 4;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 4) stages = 4;
  if (stages < 2) return 1;
  const int tmem_a_col = (bn + 31) / 32 * 32;
  if (use_ts)
    while (stages > 2 && tmem_a_col + 64 * stages > 512) --stages;
  int tmem_cols = 32;
  while (tmem_cols < (use_ts ? tmem_a_col + 64 * stages : bn)) tmem_cols <<= 1;
  const size_t smem = stages * stage_bytes + (3 * stages + 2) * sizeof(uint64_t) + 1024;
  const int k_per_split = ((nrows + splits - 1) / splits + 31) / 32 * 32;

  CUtensorMap mg, mx, mx2;
  if (use_ts ? !make_map(&mg, g, nrows, fo, ld_g, 32, kGmBM, false, true) : !make_map(&mg, g, nrows, fo, ld_g, 32, 32, true)) return 1;
  if (!make_map(&mx, x, nrows, fi, ld_x, 32, 32, true)) return 1;
  if (x2) { if (!make_map(&mx2, x2, nrows, fi, ld_x2, 32, 32, true)) return 1; }
  else mx2 = mx;
  static DeviceOnce configured;
  if (configured.need()) {
    HL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  GemmParams P;
  P.num_tiles = 0; P.tiles_n = 1; P.acc_stride = 0; P.x_tiles_each = x2 ? ntiles : 0; P.bn_part = nullptr; P.bn_nvalid = nullptr; P.tma_c = 0;
  P.M = fo; P.N = fi; P.K = nrows; P.bn = bn; P.stages = stages; P.tmem_cols = tmem_cols;
  P.bias = nullptr; P.C = reinterpret_cast<float*>(workspace); P.ldc = fi_tot; P.accumulate = 0;
  P.k_per_split = k_per_split; P.split_stride = (int64_t)fo * fi_tot; P.tmem_a_col = tmem_a_col;
  const bool fold_bias = dbias && use_ts;
  float* cs_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) +
                                          align_up((size_t)splits * (size_t)fo * (size_t)fi_tot * sizeof(float), 256));
  // two converter groups alternate over the ring stages: a group must meet every phase of the barriers it waits on, so
  // the stage count has to be even (with 3 stages group 0 would see stage 0 at rounds 0, 2, 4, ... and a parity wait
  // cannot tell round 2 from round 0)
  static int max_groups = 0;
  if (max_groups == 0) { const char* e = getenv("HL_WGRAD_GROUPS"); max_groups = e ? atoi(e) : 2; if (max_groups < 1 || max_groups > 2) max_groups = 2; }
  P.colsum_ws = fold_bias ? cs_ws : nullptr; P.conv_groups = (use_ts && max_groups == 2 && stages % 2 == 0) ? 2 : 1;
  dim3 grid(mtiles, ntiles * nx, splits);
  P.kb_first = 0;
  if (use_ts) hl::launch_pdl(gemm_tf32x3_kernel<1, 1>, grid, 64 + P.conv_groups * kGmConvThreads, smem, as_stream(stream), mg, mx, mx2, mx, P);
  else hl::launch_pdl(gemm_tf32x3_kernel<1, 0>, grid, 64 + kGmConvThreads, smem, as_stream(stream), mg, mx, mx, mx, P);
  HL_LAUNCH_CHECK("gemm_tf32x3_kernel<wgrad>");
  const int64_t n = (int64_t)fo * fi_tot;
  if (defer) {                                                       // the caller sums the splits later (hl_wgrad_reduce_batch)
    defer->partial = P.C; defer->cs_partial = cs_ws; defer->dw = dw; defer->dw2 = dw2; defer->dbias = fold_bias ? dbias : nullptr;
    defer->split_stride = P.split_stride; defer->ld_dw = ld_dw; defer->ld_dw2 = ld_dw2;
    defer->splits = splits; defer->fo = fo; defer->fi = (int32_t)fi_tot; defer->fi_first = fi;
    defer->accumulate = accumulate; defer->accumulate_bias = accumulate_bias; defer->block_start = 0; defer->reserved = 0;
    return (dbias && !fold_bias) ? 2 : HL_OK;
  }
  hl::launch_pdl(gm_split_reduce_kernel, (int)(((n + (fold_bias ? fo : 0)) * 2 + 255) / 256), 256, 0, as_stream(stream), P.C, splits, P.split_stride, fo, (int32_t)fi_tot, dw, ld_dw, accumulate, cs_ws, fold_bias ? dbias : nullptr, accumulate_bias,
      dw2, ld_dw2, fi);
  HL_LAUNCH_CHECK("gm_split_reduce_kernel");
  return (dbias && !fold_bias) ? 2 : HL_OK;
}

extern "C" int hl_wgrad_bias_tf32x3(const float* g, int64_t ld_g, const float* x, int64_t ld_x, int32_t nrows, int32_t fo,
                                    int32_t fi, float* dw, int64_t ld_dw, int accumulate, float* dbias, int accumulate_bias,
                                    void* workspace, size_t workspace_bytes, hl_stream_t stream) {
  return wgrad_launch(g, ld_g, x, ld_x, nullptr, 0, nrows, fo, fi, dw, ld_dw, nullptr, 0, accumulate, dbias, accumulate_bias,
                      workspace, workspace_bytes, stream);
}

extern "C" size_t hl_wgrad2_tf32x3_workspace(int32_t nrows, int32_t fo, int32_t fi) {
  if (nrows < 0 || fo < 1 || fi < 1) return 0;
  const int ntiles = wgrad_ntiles(fi);
  const int tiles = ((fo + hl::kGmBM - 1) / hl::kGmBM) * ntiles * 2;
  const size_t splits = (size_t)wgrad_tc_splits(nrows, tiles);
  return hl::align_up(splits * (size_t)fo * (size_t)fi * 2 * sizeof(float), 256) + 2 * splits * (size_t)fo * sizeof(float) + 256;
}

// dW1 (=|+=) g^T x1 and dW2 (=|+=) g^T x2 (x1, x2: [R, fi], the same g [R, fo]) from one launch + one reduce; dbias as above.
// Returns 1 when the shape is unsupported (caller issues two single launches).
extern "C" int hl_wgrad2_bias_tf32x3(const float* g, int64_t ld_g, const float* x1, int64_t ld_x1, const float* x2, int64_t ld_x2,
                                     int32_t nrows, int32_t fo, int32_t fi, float* dw1, int64_t ld_dw1, float* dw2,
                                     int64_t ld_dw2, int accumulate, float* dbias, int accumulate_bias, void* workspace,
                                     size_t workspace_bytes, hl_stream_t stream) {
  if (!x2 || !dw2) return HL_ERR_INVALID;
  return wgrad_launch(g, ld_g, x1, ld_x1, x2, ld_x2, nrows, fo, fi, dw1, ld_dw1, dw2, ld_dw2, accumulate, dbias, accumulate_bias,
                      workspace, workspace_bytes, stream);
}

// The weight gradient (one activation: x2 = dw2 = NULL, or two that share g) WITHOUT its split reduce: the partial planes
// stay in `workspace` (which must live until the reduce has run) and `desc` describes the reduce for
// hl_wgrad_reduce_batch.  Same return codes as hl_wgrad_bias_tf32x3; on 1 nothing was launched and desc is untouched.
extern "C" int hl_wgrad_deferred_tf32x3(const float* g, int64_t ld_g, const float* x1, int64_t ld_x1, const float* x2,
                                        int64_t ld_x2, int32_t nrows, int32_t fo, int32_t fi, float* dw1, int64_t ld_dw1,
                                        float* dw2, int64_t ld_dw2, int accumulate, float* dbias, int accumulate_bias,
                                        void* workspace, size_t workspace_bytes, hl_wgrad_reduce_desc* desc, hl_stream_t stream) {
  if (!desc || (x2 != nullptr) != (dw2 != nullptr)) return HL_ERR_INVALID;
  return wgrad_launch(g, ld_g, x1, ld_x1, x2, ld_x2, nrows, fo, fi, dw1, ld_dw1, dw2, ld_dw2, accumulate, dbias, accumulate_bias,
                      workspace, workspace_bytes, stream, desc);
}

// All deferred reduces in (n / 192 rounded up) launches; `descs` is a HOST array (copied into the kernel parameters).  Two
// descriptors of one call must not write the same elements (the caller reduces such a pair in separate calls).
extern "C" int hl_wgrad_reduce_batch(const hl_wgrad_reduce_desc* descs, int32_t n, hl_stream_t stream) {
  using namespace hl;
  if (n < 0 || (n > 0 && !descs)) return HL_ERR_INVALID;
  for (int32_t first = 0; first < n; first += kReduceBatchMax) {
    ReduceBatch B;
    B.n = n - first < kReduceBatchMax ? n - first : kReduceBatchMax;
    int64_t blocks = 0;
    for (int k = 0; k < B.n; ++k) {
      B.d[k] = descs[first + k];
      const int64_t tot4 = (((int64_t)B.d[k].fo * B.d[k].fi) >> 2) + (B.d[k].dbias ? (B.d[k].fo >> 2) : 0);
      const int64_t groups = (tot4 + 3) & ~(int64_t)3;
      if (blocks > 0x7fffffff - (groups + 31) / 32) return HL_ERR_INVALID;
      B.d[k].block_start = (int32_t)blocks;
      blocks += (groups + 31) / 32;
    }
    B.total_blocks = (int32_t)blocks;
    if (blocks == 0) continue;
    const int64_t max_grid = (int64_t)device_sm_count() * 8;
    hl::launch_pdl(gm_split_reduce_batch_kernel, (int)(blocks < max_grid ? blocks : max_grid), 256, 0, as_stream(stream), B);
    HL_LAUNCH_CHECK("gm_split_reduce_batch_kernel");
  }
  return HL_OK;
}
