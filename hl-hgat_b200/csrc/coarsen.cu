// Multi-level graph coarsening on the GPU: the greedy heavy-edge matching behind the
// reference's MLGC / MLGC_weighted (lib/Hodge_Dataset.py:241-353).  torch_cluster.graclus_cluster, which the
// reference calls, is randomised; like the CPU oracle (oracle/pyg_shim/torch_cluster) this is the deterministic
// variant: nodes are visited in id order, an unmatched node pairs with its unmatched neighbour of largest
// weight (first one in ascending neighbour order on ties), cluster id = the smaller member id.
//
// The visit order is inherently sequential per graph, so one warp owns one graph and walks its nodes; the lanes
// scan the incident-edge list of the current node in parallel (coalesced loads, one arg-max reduction per 32
// edges).  Graphs of a mini-batch are independent, so a batch of 256 CIFAR-sized graphs keeps 256 warps busy.
#include "common.cuh"

namespace hl {

__global__ void __launch_bounds__(128)
greedy_matching_kernel(const int32_t* __restrict__ node_ptr, int32_t n_graphs, const int32_t* __restrict__ inc_rowptr,
                       const int32_t* __restrict__ inc_edge, const int32_t* __restrict__ tail,
                       const int32_t* __restrict__ head, const float* __restrict__ edge_weight,
                       int32_t* __restrict__ cluster) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_graphs) return;
  const int n0 = node_ptr[g], n1 = node_ptr[g + 1];
  for (int u = n0 + lane; u < n1; u += 32) cluster[u] = -1;
  __syncwarp();
  for (int u = n0; u < n1; ++u) {
    if (cluster[u] >= 0) continue;                           // warp-uniform: same address for every lane
    const int s = inc_rowptr[u], e = inc_rowptr[u + 1];
    float best_w = -INFINITY;
    int best_v = -1, best_p = 0x7fffffff;
    for (int base = s; base < e; base += 32) {
      const int p = base + lane;
      int v = -1;
      float w = -INFINITY;
      if (p < e) {
        const int ed = inc_edge[p];
        const int t = tail[ed], h = head[ed];
        v = (t == u) ? h : t;
        if (v == u || cluster[v] >= 0) v = -1;               // self loop / already matched
        else w = edge_weight ? edge_weight[ed] : 1.f;
      }
      // arg-max over the lanes: larger weight wins, equal weights keep the earlier list position
      float cw = w;
      int cv = v, cp = v >= 0 ? p : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ow = __shfl_xor_sync(0xffffffffu, cw, off);
        const int ov = __shfl_xor_sync(0xffffffffu, cv, off);
        const int op = __shfl_xor_sync(0xffffffffu, cp, off);
        if (ov >= 0 && (cv < 0 || ow > cw || (ow == cw && op < cp))) { cw = ow; cv = ov; cp = op; }
      }
      if (cv >= 0 && (best_v < 0 || cw > best_w)) { best_w = cw; best_v = cv; best_p = cp; }   // strict >: first maximum
    }
    if (lane == 0) {
      cluster[u] = u;
      if (best_v >= 0) cluster[best_v] = u;
    }
    __syncwarp();
  }
}

}  // namespace hl

extern "C" int hl_greedy_matching(const int32_t* node_ptr, int32_t n_graphs, const int32_t* inc_rowptr,
                                  const int32_t* inc_edge, const int32_t* tail, const int32_t* head,
                                  const float* edge_weight, int32_t* cluster, hl_stream_t stream) {
  using namespace hl;
  if (n_graphs < 0) return HL_ERR_INVALID;
  if (n_graphs == 0) return HL_OK;
  if (!node_ptr || !inc_rowptr || !inc_edge || !tail || !head || !cluster) return HL_ERR_INVALID;
  const int warps_per_block = 4;
  greedy_matching_kernel<<<(n_graphs + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, as_stream(stream)>>>(
      node_ptr, n_graphs, inc_rowptr, inc_edge, tail, head, edge_weight, cluster);
  HL_LAUNCH_CHECK("greedy_matching_kernel");
  return HL_OK;
}
