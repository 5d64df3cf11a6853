// Library-level entry points: version, status strings, last CUDA error.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace hl {

thread_local char g_last_error[256] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int record_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return HL_ERR_CUDA;
}

}  // namespace hl

namespace hl { unsigned long long launches(); }
extern "C" unsigned long long hl_launch_count(void) { return hl::launches(); }

extern "C" int hl_version(void) { return HL_ABI_VERSION; }

extern "C" const char* hl_status_string(int status) {
  switch (status) {
    case HL_OK: return "ok";
    case HL_ERR_INVALID: return "invalid argument";
    case HL_ERR_WORKSPACE: return "workspace missing or too small";
    case HL_ERR_CUDA: return "CUDA error";
    case HL_ERR_ALIGN: return "misaligned pointer or leading dimension";
    default: return "unknown status";
  }
}

extern "C" const char* hl_last_cuda_error(void) { return hl::g_last_error; }

extern "C" int hl_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
