// Shared helpers for libhlhgat (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "hlhgat.h"

namespace hl {

extern thread_local char g_last_error[256];
int record_cuda_error(cudaError_t e, const char* where);
void count_launch();   // bumps the library-wide kernel launch counter (hl_launch_count)

#define HL_CUDA_CHECK(expr)                                         \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return hl::record_cuda_error(_e, #expr); \
  } while (0)

#define HL_LAUNCH_CHECK(name)                                      \
  do {                                                             \
    hl::count_launch();                                            \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return hl::record_cuda_error(_e, name); \
  } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may start while its stream predecessor is
// still running -- as soon as every CTA of the predecessor has executed pdl_trigger() or exited -- and must execute
// pdl_wait() before it touches anything the predecessor wrote (the wait returns once the predecessor has completed and its
// writes are visible; without the attribute both instructions are no-ops).  What overlaps is the launch latency and the
// dependent's prologue (barrier init, TMEM allocation) with the predecessor's tail; stream capture records the edge as a
// programmatic one.  HL_PDL=0 launches everything without the attribute.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HL_PDL"); v = e ? atoi(e) : 1; }
  return v != 0;
}
template <typename... Exp, typename... Act>
inline cudaError_t launch_pdl(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
#ifndef HL_PDL_EARLY_TRIGGER
#define HL_PDL_EARLY_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if HL_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// for the kernels with a long tail after their last dependent-visible decision (GEMM: the epilogue): dependents may be
// scheduled from here (HL_PDL_LATE_TRIGGER=0 at compile time: only when the grid has exited)
#ifndef HL_PDL_LATE_TRIGGER
#define HL_PDL_LATE_TRIGGER 1
#endif
__device__ __forceinline__ void pdl_trigger_late() {
#if HL_PDL_LATE_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One-time PER-DEVICE configuration (cudaFuncSetAttribute and the SM count belong to a device, not to the process):
// `need()` is true the first time it is called with a given device current.
struct DeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// multiprocessor count of the CURRENT device (cached per device)
static inline int device_sm_count() {
  static int cached[64] = {};
  int d = 0, n = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 148;
  if (cached[d] > 0) return cached[d];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
  cached[d] = n;
  return n;
}

static inline cudaStream_t as_stream(hl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// widest vector (in floats) usable for a [rows, width] slice with leading dimension ld
static inline int vec_for(const void* p, int64_t ld, int32_t width, int want) {
  if (p == nullptr) return want;
  int v = want;
  while (v > 1 && (width % v != 0 || ld % v != 0 || !aligned_to(p, sizeof(float) * v))) v >>= 1;
  return v;
}

template <int V> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int V>
struct Pack {
  float v[V];
};

template <int V>
__device__ __forceinline__ Pack<V> ld_pack(const float* p) {
  Pack<V> r;
  if constexpr (V == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

// plain (coherent) load: for operands that may alias the output buffer
template <int V>
__device__ __forceinline__ Pack<V> ld_pack_coherent(const float* p) {
  Pack<V> r;
  if constexpr (V == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}

template <int V>
__device__ __forceinline__ void st_pack(float* p, const Pack<V>& r) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// lanes per row group: smallest power of two >= ceil(width / V), capped at 32
static inline int group_lanes(int32_t width, int v) {
  int chunks = (width + v - 1) / v;
  int g = 1;
  while (g < chunks && g < 32) g <<= 1;
  return g;
}

}  // namespace hl
