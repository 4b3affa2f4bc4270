// Host-side helpers shared by the kernel launchers: per-DEVICE one-time kernel attribute opt-in and SM count.
// cudaFuncSetAttribute applies to the current device only, so a process that drives several GPUs (one engine per
// device, diffusion_pullback_b200.api._engine_for) must opt in once per (device, kernel), not once per process.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <set>
#include <utility>

namespace pbhost {

constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// Opt the kernel `func` in to `bytes` of dynamic shared memory on the current device (first launch there only).
template <class F>
inline const char* optin_smem(F* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  const std::pair<int, const void*> key{current_device(), reinterpret_cast<const void*>(func)};
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(key)) return nullptr;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  done.insert(key);
  return nullptr;
}

// SM count of the current device (148 on B200), cached per device.
inline int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  if (!n[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

}  // namespace pbhost
