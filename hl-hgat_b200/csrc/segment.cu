// Segmented reductions and gathers behind the node<->edge simplex transfer, the attention gate,
// cluster pooling and the per-graph readout (hl_segment_reduce, hl_endpoint_gather, hl_owner_gather,
// hl_att_gate_fwd/bwd).  All HBM-bound; no atomics anywhere: each output row is owned by one lane
// group that walks its CSR bucket in ascending order, so results are deterministic run to run.
#include <cstdlib>

#include "common.cuh"

namespace hl {

constexpr int kSegThreads = 256;

// ---------------------------------------------------------------------------------------------
// dst[r,:] = post( sum_p pre(src[m_p,:]) )
// ---------------------------------------------------------------------------------------------
template <int V, int CH>
__global__ void __launch_bounds__(kSegThreads)
segment_reduce_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t nrows,
                      const float* __restrict__ src, int64_t ld_src, const float* __restrict__ src_scale,
                      float* __restrict__ dst, int64_t ld_dst, int32_t width, int32_t G, int post,
                      const float* __restrict__ row_scale, float cscale) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int rows_per_block = kSegThreads / G;
  const int gl = threadIdx.x & (G - 1);
  const int row = blockIdx.x * rows_per_block + (int)(threadIdx.x / G);
  if (row >= nrows) return;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  const int col0 = blockIdx.y * (G * V * CH) + gl * V;
  bool act[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) act[ch] = (col0 + ch * G * V) < width;
  Pack<V> acc[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[ch].v[i] = 0.f;

  const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int base = start; base < end; base += G) {
    const int p = base + gl;
    int m = 0;
    float s = 1.f;
    if (p < end) {
      m = colidx ? __ldg(colidx + p) : p;
      if (src_scale) s = __ldg(src_scale + m);
    }
    const int cnt = min(G, end - base);
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
      int mj[4];
      float sj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mj[u] = __shfl_sync(gmask, m, j + u, G);
        sj[u] = __shfl_sync(gmask, s, j + u, G);
      }
      Pack<V> x[4][CH];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (act[ch]) x[u][ch] = ld_pack<V>(src + (int64_t)mj[u] * ld_src + col0 + ch * G * V);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (act[ch])
#pragma unroll
            for (int i = 0; i < V; ++i) {
              const float t = src_scale ? __fmul_rn(x[u][ch].v[i], sj[u]) : x[u][ch].v[i];
              acc[ch].v[i] = __fadd_rn(acc[ch].v[i], t);
            }
    }
    for (; j < cnt; ++j) {
      const int mj = __shfl_sync(gmask, m, j, G);
      const float sj = __shfl_sync(gmask, s, j, G);
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        if (act[ch]) {
          Pack<V> x = ld_pack<V>(src + (int64_t)mj * ld_src + col0 + ch * G * V);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float t = src_scale ? __fmul_rn(x.v[i], sj) : x.v[i];
            acc[ch].v[i] = __fadd_rn(acc[ch].v[i], t);
          }
        }
    }
  }

  float mul = 1.f, div = 1.f;
  if (post == HL_POST_CONST) mul = cscale;
  else if (post == HL_POST_RCP_ROW) mul = __fdiv_rn(1.f, __ldg(row_scale + row));
  else if (post == HL_POST_MEAN) div = (float)max(end - start, 1);
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    if (!act[ch]) continue;
    Pack<V> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t = acc[ch].v[i];
      if (post == HL_POST_MEAN) t = __fdiv_rn(t, div);
      else if (post != HL_POST_NONE) t = __fmul_rn(mul, t);
      o.v[i] = t;
    }
    st_pack<V>(dst + (int64_t)row * ld_dst + col0 + ch * G * V, o);
  }
}

// The same reduction for SHORT buckets (the |B1| transfers: ~2 incident edges per node on molecular graphs, a handful on
// kNN graphs).  With one lane group per row the kernel is a chain of three dependent global loads (rowptr -> colidx ->
// source rows) per 2-3 gathered rows: latency-bound at ~30 % of the HBM roofline.  Here a lane group owns RB consecutive
// rows: ONE coalesced read fetches their RB + 1 row pointers (and per-row scales), ONE read the member indices of all RB
// buckets, and the source rows of up to four members are in flight together, whichever bucket they belong to.  Members
// are still added in ascending bucket order into one running accumulator that is flushed at every bucket boundary, so
// every output row sees exactly the additions of the per-row kernel, in the same order: bit-identical.
template <int V, int CH, int RB>
__global__ void __launch_bounds__(kSegThreads)
segment_reduce_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t nrows,
                           const float* __restrict__ src, int64_t ld_src, const float* __restrict__ src_scale,
                           float* __restrict__ dst, int64_t ld_dst, int32_t width, int32_t G, int post,
                           const float* __restrict__ row_scale, float cscale) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int groups_per_block = kSegThreads / G;
  const int gl = threadIdx.x & (G - 1);
  const int r0 = (blockIdx.x * groups_per_block + (int)(threadIdx.x / G)) * RB;
  if (r0 >= nrows) return;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  const int col0 = blockIdx.y * (G * V * CH) + gl * V;
  bool act[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) act[ch] = (col0 + ch * G * V) < width;
  const int nr = min(RB, nrows - r0);                                  // rows this group really owns
  // lane k <= RB holds rowptr[r0 + k] (clamped) and, for k < RB, the per-row scale
  const int myptr = __ldg(rowptr + min(r0 + min(gl, RB), nrows));
  float myscale = 1.f;
  if (post == HL_POST_RCP_ROW && gl < nr) myscale = __ldg(row_scale + r0 + gl);
  const int first = __shfl_sync(gmask, myptr, 0, G), last = __shfl_sync(gmask, myptr, nr, G);

  Pack<V> acc[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[ch].v[i] = 0.f;
  int k = 0;                                                           // bucket being accumulated
  int kstart = first, kend = __shfl_sync(gmask, myptr, 1, G);

  auto flush = [&]() {                                                 // write row r0 + k, reset the accumulator, advance
    float mul = 1.f, div = 1.f;
    if (post == HL_POST_CONST) mul = cscale;
    else if (post == HL_POST_RCP_ROW) mul = __fdiv_rn(1.f, __shfl_sync(gmask, myscale, k, G));
    else if (post == HL_POST_MEAN) div = (float)max(kend - kstart, 1);
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      if (act[ch]) {
        Pack<V> o;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float t = acc[ch].v[i];
          if (post == HL_POST_MEAN) t = __fdiv_rn(t, div);
          else if (post != HL_POST_NONE) t = __fmul_rn(mul, t);
          o.v[i] = t;
          acc[ch].v[i] = 0.f;
        }
        st_pack<V>(dst + (int64_t)(r0 + k) * ld_dst + col0 + ch * G * V, o);
      }
    }
    ++k;
    kstart = kend;
    kend = __shfl_sync(gmask, myptr, min(k + 1, RB), G);
  };

  for (int base = first; base < last; base += G) {
    const int p = base + gl;
    int m = 0;
    float s = 1.f;
    if (p < last) {
      m = colidx ? __ldg(colidx + p) : p;
      if (src_scale) s = __ldg(src_scale + m);
    }
    const int cnt = min(G, last - base);
    for (int j = 0; j < cnt; j += 4) {
      int mj[4];
      float sj[4];
      Pack<V> x[4][CH];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mj[u] = __shfl_sync(gmask, m, min(j + u, G - 1), G);
        sj[u] = __shfl_sync(gmask, s, min(j + u, G - 1), G);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j + u < cnt)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (act[ch]) x[u][ch] = ld_pack<V>(src + (int64_t)mj[u] * ld_src + col0 + ch * G * V);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u >= cnt) break;
        while (base + j + u >= kend) flush();                        // bucket boundary (also skips empty buckets)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (act[ch])
#pragma unroll
            for (int i = 0; i < V; ++i) {
              const float t = src_scale ? __fmul_rn(x[u][ch].v[i], sj[u]) : x[u][ch].v[i];
              acc[ch].v[i] = __fadd_rn(acc[ch].v[i], t);
            }
      }
    }
  }
  while (k < nr) flush();                                              // the last bucket and any empty ones after it
}

// ---------------------------------------------------------------------------------------------
// dst[e,:] = cscale * (f(src[tail[e]]) + f(src[head[e]]))
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
endpoint_gather_kernel(const int32_t* __restrict__ tail, const int32_t* __restrict__ head, int32_t nedges,
                       const float* __restrict__ src, int64_t ld_src, const float* __restrict__ node_rcp,
                       float* __restrict__ dst, int64_t ld_dst, int32_t chunks, float cscale) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int64_t total = (int64_t)nedges * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx / chunks);
    const int col = (int)(idx - (int64_t)e * chunks) * V;
    const int t = __ldg(tail + e), h = __ldg(head + e);
    Pack<V> a = ld_pack<V>(src + (int64_t)t * ld_src + col);
    Pack<V> b = ld_pack<V>(src + (int64_t)h * ld_src + col);
    Pack<V> o;
    if (node_rcp) {
      const float rt = __fdiv_rn(1.f, __ldg(node_rcp + t)), rh = __fdiv_rn(1.f, __ldg(node_rcp + h));
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = __fmul_rn(cscale, __fadd_rn(__fmul_rn(rt, a.v[i]), __fmul_rn(rh, b.v[i])));
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = __fmul_rn(cscale, __fadd_rn(a.v[i], b.v[i]));
    }
    st_pack<V>(dst + (int64_t)e * ld_dst + col, o);
  }
}

// ---------------------------------------------------------------------------------------------
// dst[e,:] = cscale * | src[head[e],:] - src[tail[e],:] |      (signed boundary B1^T, then abs)
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
boundary_absdiff_fwd_kernel(const int32_t* __restrict__ tail, const int32_t* __restrict__ head, int32_t nedges,
                            const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst,
                            int32_t chunks, float cscale) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int64_t total = (int64_t)nedges * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx / chunks);
    const int col = (int)(idx - (int64_t)e * chunks) * V;
    Pack<V> a = ld_pack<V>(src + (int64_t)__ldg(tail + e) * ld_src + col);
    Pack<V> b = ld_pack<V>(src + (int64_t)__ldg(head + e) * ld_src + col);
    Pack<V> o;
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = __fmul_rn(cscale, fabsf(__fadd_rn(b.v[i], -a.v[i])));
    st_pack<V>(dst + (int64_t)e * ld_dst + col, o);
  }
}

// dsrc[n,:] = cscale * sum_{e incident to n, ascending e} s(n,e) * sgn(src[head_e]-src[tail_e]) * g[e,:],
// s(n,e) = +1 if n is the head of e, -1 if it is the tail.  One lane group per node; no atomics.
template <int V>
__global__ void __launch_bounds__(256)
boundary_absdiff_bwd_kernel(const int32_t* __restrict__ inc_rowptr, const int32_t* __restrict__ inc_edge,
                            const int32_t* __restrict__ tail, const int32_t* __restrict__ head, int32_t nnodes,
                            const float* __restrict__ src, int64_t ld_src, const float* __restrict__ g, int64_t ld_g,
                            float* __restrict__ dsrc, int64_t ld_dsrc, int32_t width, int32_t G, float cscale) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int rows_per_block = 256 / G;
  const int gl = threadIdx.x & (G - 1);
  const int n = blockIdx.x * rows_per_block + (int)(threadIdx.x / G);
  if (n >= nnodes) return;
  const int start = __ldg(inc_rowptr + n), end = __ldg(inc_rowptr + n + 1);
  for (int col = gl * V; col < width; col += G * V) {
    Pack<V> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
    for (int p = start; p < end; ++p) {
      const int e = __ldg(inc_edge + p);
      const int t = __ldg(tail + e), h = __ldg(head + e);
      if (t == h) continue;                                   // self-pair (ghost padding): B1 column is zero
      const float s = (h == n) ? 1.f : -1.f;
      Pack<V> a = ld_pack<V>(src + (int64_t)t * ld_src + col);
      Pack<V> b = ld_pack<V>(src + (int64_t)h * ld_src + col);
      Pack<V> gv = ld_pack<V>(g + (int64_t)e * ld_g + col);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = __fadd_rn(b.v[i], -a.v[i]);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        acc.v[i] = __fadd_rn(acc.v[i], s * sg * gv.v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = __fmul_rn(cscale, acc.v[i]);
    st_pack<V>(dsrc + (int64_t)n * ld_dsrc + col, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// dsrc[m,:] = g[owner[m],:] * w_m * scale_m ; dscale[m] = w_m <g[owner[m],:], src[m,:]>
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256)
owner_gather_kernel(const int32_t* __restrict__ owner, int32_t nsrc, const float* __restrict__ g, int64_t ld_g,
                    const float* __restrict__ owner_scale, const float* __restrict__ src_scale,
                    const float* __restrict__ src, int64_t ld_src, float* __restrict__ dsrc, int64_t ld_dsrc,
                    float* __restrict__ dscale, int32_t width, int32_t G) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int rows_per_block = 256 / G;
  const int gl = threadIdx.x & (G - 1);
  const int m = blockIdx.x * rows_per_block + (int)(threadIdx.x / G);
  if (m >= nsrc) return;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  const int o = __ldg(owner + m);
  const float w = (o >= 0 && owner_scale) ? __ldg(owner_scale + o) : 1.f;
  const float sc = src_scale ? __ldg(src_scale + m) : 1.f;
  float dot = 0.f;
  for (int col = gl * V; col < width; col += G * V) {
    Pack<V> r;
    if (o >= 0) {
      Pack<V> gv = ld_pack<V>(g + (int64_t)o * ld_g + col);
      if (dscale) {
        Pack<V> sv = ld_pack<V>(src + (int64_t)m * ld_src + col);
#pragma unroll
        for (int i = 0; i < V; ++i) dot = fmaf(gv.v[i], sv.v[i], dot);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) r.v[i] = gv.v[i] * (w * sc);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) r.v[i] = 0.f;
    }
    if (dsrc) st_pack<V>(dsrc + (int64_t)m * ld_dsrc + col, r);
  }
  if (dscale) {
    for (int off = G >> 1; off > 0; off >>= 1) dot += __shfl_xor_sync(gmask, dot, off, G);
    if (gl == 0) dscale[m] = dot * w;
  }
}

// ---------------------------------------------------------------------------------------------
// attention gate
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gate_act(float pre, int sigma) {
  return sigma == HL_SIGMA_SIGMOID ? 1.f / (1.f + expf(-pre)) : fmaxf(pre, 0.f);
}

template <int V>
__global__ void __launch_bounds__(256)
att_gate_fwd_kernel(const float* __restrict__ qc, const float* __restrict__ qs, const float* __restrict__ k,
                    int32_t nrows, int32_t dk, int32_t G, float lambda, float inv_sqrt_dk, int sigma,
                    float* __restrict__ a) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int rows_per_block = 256 / G;
  const int gl = threadIdx.x & (G - 1);
  const int r = blockIdx.x * rows_per_block + (int)(threadIdx.x / G);
  if (r >= nrows) return;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  float dc = 0.f, ds = 0.f;
  for (int col = gl * V; col < dk; col += G * V) {
    Pack<V> kv = ld_pack<V>(k + (int64_t)r * dk + col);
    Pack<V> c = ld_pack<V>(qc + (int64_t)r * dk + col);
    Pack<V> s = ld_pack<V>(qs + (int64_t)r * dk + col);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      dc = fmaf(c.v[i], kv.v[i], dc);
      ds = fmaf(s.v[i], kv.v[i], ds);
    }
  }
  for (int off = G >> 1; off > 0; off >>= 1) {
    dc += __shfl_xor_sync(gmask, dc, off, G);
    ds += __shfl_xor_sync(gmask, ds, off, G);
  }
  if (gl == 0) a[r] = gate_act(((1.f - lambda) * dc + lambda * ds) * inv_sqrt_dk, sigma);
}

template <int V>
__global__ void __launch_bounds__(256)
att_gate_bwd_kernel(const float* __restrict__ qc, const float* __restrict__ qs, const float* __restrict__ k,
                    const float* __restrict__ a, const float* __restrict__ da, int32_t nrows, int32_t dk,
                    float lambda, float inv_sqrt_dk, int sigma,
                    float* __restrict__ dqc, float* __restrict__ dqs, float* __restrict__ dkk) {
  hl::pdl_trigger();
  hl::pdl_wait();   // programmatic dependent launch: see common.cuh
  const int chunks = dk / V;
  const int64_t total = (int64_t)nrows * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / chunks);
    const int col = (int)(idx - (int64_t)r * chunks) * V;
    const float av = __ldg(a + r);
    const float dact = sigma == HL_SIGMA_SIGMOID ? av * (1.f - av) : (av > 0.f ? 1.f : 0.f);
    const float dpre = __ldg(da + r) * dact * inv_sqrt_dk;
    const float wc = dpre * (1.f - lambda), ws = dpre * lambda;
    Pack<V> kv = ld_pack<V>(k + (int64_t)r * dk + col);
    Pack<V> c = ld_pack<V>(qc + (int64_t)r * dk + col);
    Pack<V> s = ld_pack<V>(qs + (int64_t)r * dk + col);
    Pack<V> oc, os, ok;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      oc.v[i] = wc * kv.v[i];
      os.v[i] = ws * kv.v[i];
      ok.v[i] = wc * c.v[i] + ws * s.v[i];
    }
    st_pack<V>(dqc + (int64_t)r * dk + col, oc);
    st_pack<V>(dqs + (int64_t)r * dk + col, os);
    st_pack<V>(dkk + (int64_t)r * dk + col, ok);
  }
}

static int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148LL * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace hl

extern "C" int hl_segment_reduce(const int32_t* rowptr, const int32_t* colidx, int32_t nrows,
                                 const float* src, int64_t ld_src, const float* src_scale,
                                 float* dst, int64_t ld_dst, int32_t width,
                                 int post, const float* row_scale, float cscale, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 0 || width < 1 || post < HL_POST_NONE || post > HL_POST_MEAN) return HL_ERR_INVALID;
  if (nrows == 0) return HL_OK;
  if (!rowptr || !src || !dst || (post == HL_POST_RCP_ROW && !row_scale)) return HL_ERR_INVALID;
  int V = vec_for(src, ld_src, width, 4);
  V = min(V, vec_for(dst, ld_dst, width, V));
  const int G = group_lanes(width, V);
  const int chunks = (width + V - 1) / V;
  int CH = chunks > G ? 2 : 1;
  // rows of 257-384 floats (the 320-wide dense-connection buffers the attention pooling averages): three chunks per lane keep
  // the row in ONE lane group instead of a second, mostly idle column tile that repeats all the index work
  static int use_rows = -1;
  if (use_rows < 0) { const char* e = getenv("HL_SEG_ROWS"); use_rows = e ? atoi(e) : 8; }
  const bool rows_variant = use_rows && G >= 16 && V == 4 && nrows >= 4096;
  if (rows_variant && use_rows >= 8 && chunks > 2 * G && chunks <= 3 * G) CH = 3;
  const int tile_w = G * V * CH;
  dim3 grid((nrows + kSegThreads / G - 1) / (kSegThreads / G), (width + tile_w - 1) / tile_w);
  // row-batched variant (8 rows per lane group; HL_SEG_ROWS=4: 4 rows) whenever a group has the lanes to hold the row
  // pointers; HL_SEG_ROWS=0 selects the per-row kernel for A/B runs.  Measured on the x16 ZINC incidence stack (B200):
  // s2t at F=64 1.87 -> 2.94 TB/s, at F=256 1.97 -> 3.80 TB/s; adjoint of t2s 1.96 -> 3.28 / 2.02 -> 3.98 TB/s
  if (rows_variant) {
    const int RB = use_rows >= 8 ? 8 : 4;
    const int groups = (nrows + RB - 1) / RB;
    dim3 grid_rb((groups + kSegThreads / G - 1) / (kSegThreads / G), (width + tile_w - 1) / tile_w);
#define HL_SEGR_CASE(CC, RR)                                                                                   \
    hl::launch_pdl(segment_reduce_rows_kernel<4, CC, RR>, grid_rb, kSegThreads, 0, as_stream(stream), \
        rowptr, colidx, nrows, src, ld_src, src_scale, dst, ld_dst, width, G, post, row_scale, cscale)
    if (RB == 8) { if (CH == 3) HL_SEGR_CASE(3, 8); else if (CH == 2) HL_SEGR_CASE(2, 8); else HL_SEGR_CASE(1, 8); }
    else { if (CH == 2) HL_SEGR_CASE(2, 4); else HL_SEGR_CASE(1, 4); }
#undef HL_SEGR_CASE
    HL_LAUNCH_CHECK("segment_reduce_rows_kernel");
    return HL_OK;
  }
#define HL_SEG_CASE(VV, CC)                                                                               \
  hl::launch_pdl(segment_reduce_kernel<VV, CC>, grid, kSegThreads, 0, as_stream(stream), \
      rowptr, colidx, nrows, src, ld_src, src_scale, dst, ld_dst, width, G, post, row_scale, cscale)
  if (V == 4) { if (CH == 2) HL_SEG_CASE(4, 2); else HL_SEG_CASE(4, 1); }
  else if (V == 2) { if (CH == 2) HL_SEG_CASE(2, 2); else HL_SEG_CASE(2, 1); }
  else { if (CH == 2) HL_SEG_CASE(1, 2); else HL_SEG_CASE(1, 1); }
#undef HL_SEG_CASE
  HL_LAUNCH_CHECK("segment_reduce_kernel");
  return HL_OK;
}

extern "C" int hl_endpoint_gather(const int32_t* tail, const int32_t* head, int32_t nedges,
                                  const float* src, int64_t ld_src, const float* node_rcp,
                                  float* dst, int64_t ld_dst, int32_t width, float cscale, hl_stream_t stream) {
  using namespace hl;
  if (nedges < 0 || width < 1) return HL_ERR_INVALID;
  if (nedges == 0) return HL_OK;
  if (!tail || !head || !src || !dst) return HL_ERR_INVALID;
  int V = vec_for(src, ld_src, width, 4);
  V = min(V, vec_for(dst, ld_dst, width, V));
  const int chunks = width / V;
  const int grid = grid_for((int64_t)nedges * chunks, 256);
  if (V == 4) hl::launch_pdl(endpoint_gather_kernel<4>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, node_rcp, dst, ld_dst, chunks, cscale);
  else if (V == 2) hl::launch_pdl(endpoint_gather_kernel<2>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, node_rcp, dst, ld_dst, chunks, cscale);
  else hl::launch_pdl(endpoint_gather_kernel<1>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, node_rcp, dst, ld_dst, chunks, cscale);
  HL_LAUNCH_CHECK("endpoint_gather_kernel");
  return HL_OK;
}

extern "C" int hl_boundary_absdiff_fwd(const int32_t* tail, const int32_t* head, int32_t nedges,
                                       const float* src, int64_t ld_src, float* dst, int64_t ld_dst,
                                       int32_t width, float cscale, hl_stream_t stream) {
  using namespace hl;
  if (nedges < 0 || width < 1) return HL_ERR_INVALID;
  if (nedges == 0) return HL_OK;
  if (!tail || !head || !src || !dst) return HL_ERR_INVALID;
  int V = vec_for(src, ld_src, width, 4);
  V = min(V, vec_for(dst, ld_dst, width, V));
  const int chunks = width / V;
  const int grid = grid_for((int64_t)nedges * chunks, 256);
  if (V == 4) hl::launch_pdl(boundary_absdiff_fwd_kernel<4>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, dst, ld_dst, chunks, cscale);
  else if (V == 2) hl::launch_pdl(boundary_absdiff_fwd_kernel<2>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, dst, ld_dst, chunks, cscale);
  else hl::launch_pdl(boundary_absdiff_fwd_kernel<1>, grid, 256, 0, as_stream(stream), tail, head, nedges, src, ld_src, dst, ld_dst, chunks, cscale);
  HL_LAUNCH_CHECK("boundary_absdiff_fwd_kernel");
  return HL_OK;
}

extern "C" int hl_boundary_absdiff_bwd(const int32_t* inc_rowptr, const int32_t* inc_edge,
                                       const int32_t* tail, const int32_t* head, int32_t nnodes,
                                       const float* src, int64_t ld_src, const float* g, int64_t ld_g,
                                       float* dsrc, int64_t ld_dsrc, int32_t width, float cscale, hl_stream_t stream) {
  using namespace hl;
  if (nnodes < 0 || width < 1) return HL_ERR_INVALID;
  if (nnodes == 0) return HL_OK;
  if (!inc_rowptr || !inc_edge || !tail || !head || !src || !g || !dsrc) return HL_ERR_INVALID;
  int V = vec_for(src, ld_src, width, 4);
  V = min(V, vec_for(g, ld_g, width, V));
  V = min(V, vec_for(dsrc, ld_dsrc, width, V));
  const int G = group_lanes(width, V);
  const int grid = (nnodes + 256 / G - 1) / (256 / G);
  if (V == 4) hl::launch_pdl(boundary_absdiff_bwd_kernel<4>, grid, 256, 0, as_stream(stream), inc_rowptr, inc_edge, tail, head, nnodes, src, ld_src, g, ld_g, dsrc, ld_dsrc, width, G, cscale);
  else if (V == 2) hl::launch_pdl(boundary_absdiff_bwd_kernel<2>, grid, 256, 0, as_stream(stream), inc_rowptr, inc_edge, tail, head, nnodes, src, ld_src, g, ld_g, dsrc, ld_dsrc, width, G, cscale);
  else hl::launch_pdl(boundary_absdiff_bwd_kernel<1>, grid, 256, 0, as_stream(stream), inc_rowptr, inc_edge, tail, head, nnodes, src, ld_src, g, ld_g, dsrc, ld_dsrc, width, G, cscale);
  HL_LAUNCH_CHECK("boundary_absdiff_bwd_kernel");
  return HL_OK;
}

extern "C" int hl_owner_gather(const int32_t* owner, int32_t nsrc, const float* g, int64_t ld_g,
                               const float* owner_scale, const float* src_scale,
                               const float* src, int64_t ld_src,
                               float* dsrc, int64_t ld_dsrc, float* dscale, int32_t width, hl_stream_t stream) {
  using namespace hl;
  if (nsrc < 0 || width < 1) return HL_ERR_INVALID;
  if (nsrc == 0) return HL_OK;
  if (!owner || !g || (!dsrc && !dscale) || (dscale && !src)) return HL_ERR_INVALID;
  int V = vec_for(g, ld_g, width, 4);
  V = min(V, vec_for(dsrc, ld_dsrc, width, V));
  V = min(V, vec_for(dscale ? src : nullptr, ld_src, width, V));
  const int G = group_lanes(width, V);
  const int grid = (nsrc + 256 / G - 1) / (256 / G);
  if (V == 4) hl::launch_pdl(owner_gather_kernel<4>, grid, 256, 0, as_stream(stream), owner, nsrc, g, ld_g, owner_scale, src_scale, src, ld_src, dsrc, ld_dsrc, dscale, width, G);
  else if (V == 2) hl::launch_pdl(owner_gather_kernel<2>, grid, 256, 0, as_stream(stream), owner, nsrc, g, ld_g, owner_scale, src_scale, src, ld_src, dsrc, ld_dsrc, dscale, width, G);
  else hl::launch_pdl(owner_gather_kernel<1>, grid, 256, 0, as_stream(stream), owner, nsrc, g, ld_g, owner_scale, src_scale, src, ld_src, dsrc, ld_dsrc, dscale, width, G);
  HL_LAUNCH_CHECK("owner_gather_kernel");
  return HL_OK;
}

extern "C" int hl_att_gate_fwd(const float* qc, const float* qs, const float* k, int32_t nrows, int32_t dk,
                               float lambda, int sigma, float* a, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 0 || dk < 1 || (sigma != HL_SIGMA_SIGMOID && sigma != HL_SIGMA_RELU)) return HL_ERR_INVALID;
  if (nrows == 0) return HL_OK;
  if (!qc || !qs || !k || !a) return HL_ERR_INVALID;
  int V = vec_for(qc, dk, dk, 4);
  V = min(V, vec_for(qs, dk, dk, V));
  V = min(V, vec_for(k, dk, dk, V));
  const int G = group_lanes(dk, V);
  const int grid = (nrows + 256 / G - 1) / (256 / G);
  const float isd = 1.f / sqrtf((float)dk);
  if (V == 4) hl::launch_pdl(att_gate_fwd_kernel<4>, grid, 256, 0, as_stream(stream), qc, qs, k, nrows, dk, G, lambda, isd, sigma, a);
  else if (V == 2) hl::launch_pdl(att_gate_fwd_kernel<2>, grid, 256, 0, as_stream(stream), qc, qs, k, nrows, dk, G, lambda, isd, sigma, a);
  else hl::launch_pdl(att_gate_fwd_kernel<1>, grid, 256, 0, as_stream(stream), qc, qs, k, nrows, dk, G, lambda, isd, sigma, a);
  HL_LAUNCH_CHECK("att_gate_fwd_kernel");
  return HL_OK;
}

extern "C" int hl_att_gate_bwd(const float* qc, const float* qs, const float* k, const float* a, const float* da,
                               int32_t nrows, int32_t dk, float lambda, int sigma,
                               float* dqc, float* dqs, float* dkk, hl_stream_t stream) {
  using namespace hl;
  if (nrows < 0 || dk < 1 || (sigma != HL_SIGMA_SIGMOID && sigma != HL_SIGMA_RELU)) return HL_ERR_INVALID;
  if (nrows == 0) return HL_OK;
  if (!qc || !qs || !k || !a || !da || !dqc || !dqs || !dkk) return HL_ERR_INVALID;
  int V = 4;
  const void* ps[6] = {qc, qs, k, dqc, dqs, dkk};
  for (auto p : ps) V = min(V, vec_for(p, dk, dk, V));
  const int grid = grid_for((int64_t)nrows * (dk / V), 256);
  const float isd = 1.f / sqrtf((float)dk);
  if (V == 4) hl::launch_pdl(att_gate_bwd_kernel<4>, grid, 256, 0, as_stream(stream), qc, qs, k, a, da, nrows, dk, lambda, isd, sigma, dqc, dqs, dkk);
  else if (V == 2) hl::launch_pdl(att_gate_bwd_kernel<2>, grid, 256, 0, as_stream(stream), qc, qs, k, a, da, nrows, dk, lambda, isd, sigma, dqc, dqs, dkk);
  else hl::launch_pdl(att_gate_bwd_kernel<1>, grid, 256, 0, as_stream(stream), qc, qs, k, a, da, nrows, dk, lambda, isd, sigma, dqc, dqs, dkk);
  HL_LAUNCH_CHECK("att_gate_bwd_kernel");
  return HL_OK;
}
