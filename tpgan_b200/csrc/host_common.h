// Host-side helpers shared by the translation units of libtpgan_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace tpg {
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;
int device_sm_count();
extern std::atomic<int> g_deterministic;   // tpgan_set_deterministic
int* device_status_word();   // device alias of the host-visible status word (tpgan_kernel_status), nullptr if unavailable

#define TPG_CHECK_LAUNCH(name)                                                                        \
  do {                                                                                                \
    cudaError_t e__ = cudaGetLastError();                                                             \
    if (e__ != cudaSuccess) return tpg::set_error(-2, "%s launch: %s", name, cudaGetErrorString(e__)); \
    tpg::g_launches.fetch_add(1, std::memory_order_relaxed);                                          \
  } while (0)
}  // namespace tpg
