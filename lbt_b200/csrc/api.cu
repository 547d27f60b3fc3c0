// Library-level entry points of liblbt_b200: version, error strings, device check, launch counter.
#include "common.cuh"

#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

namespace lbt {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_pdl{1};
std::atomic<int> g_carveout{-2};   // -2: not initialised (reads LBT_CARVEOUT on first use)

void apply_carveout(const void* kernel) {
  int c = g_carveout.load(std::memory_order_relaxed);
  if (c == -2) {
    const char* e = std::getenv("LBT_CARVEOUT");
    c = e ? std::atoi(e) : -1;
    g_carveout.store(c, std::memory_order_relaxed);
  }
  if (c < 0) return;
  static std::mutex mu;
  static std::unordered_map<const void*, int> done;
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find(kernel);
  if (it != done.end() && it->second == c) return;
  (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c > 100 ? 100 : c);
  (void)cudaGetLastError();
  done[kernel] = c;
}

namespace {
thread_local std::string t_last_error;
std::once_flag g_once[16];
DeviceInfo g_info[16];
}  // namespace

void set_cuda_error(cudaError_t e, const char* where) {
  t_last_error = std::string(where ? where : "?") + ": " + cudaGetErrorName(e) + " — " + cudaGetErrorString(e);
}

const DeviceInfo& device_info() {
  static DeviceInfo bad;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 16) {
    if (e != cudaSuccess) set_cuda_error(e, "cudaGetDevice");
    (void)cudaGetLastError();
    return bad;
  }
  std::call_once(g_once[dev], [dev]() {
    DeviceInfo d;
    d.device = dev;
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    d.ok = (d.cc_major == 10 && d.cc_minor == 0);  // sm_100a cubin only: B200
    g_info[dev] = d;
  });
  return g_info[dev];
}

}  // namespace lbt

extern "C" int lbt_version(void) { return LBT_VERSION; }

extern "C" const char* lbt_strerror(int status) {
  switch (status) {
    case LBT_OK: return "ok";
    case LBT_EINVAL: return "invalid argument";
    case LBT_EUNSUPPORTED: return "unsupported shape, alignment or kind";
    case LBT_EARCH: return "device is not sm_100 (B200); this library carries sm_100a code only";
    case LBT_ECUDA: return "CUDA call failed (see lbt_last_cuda_error)";
    case LBT_EWORKSPACE: return "workspace too small";
    default: return "unknown lbt status";
  }
}

extern "C" const char* lbt_last_cuda_error(void) { return lbt::t_last_error.c_str(); }

extern "C" uint64_t lbt_launch_count(void) { return lbt::g_launches.load(std::memory_order_relaxed); }

// Bench / test knob (not in lbt.h): shared-memory carveout percent for all launch_pdl kernels, -1 = driver default.
extern "C" int lbt_set_carveout(int percent) {
  lbt::g_carveout.store(percent < 0 ? -1 : percent, std::memory_order_relaxed);
  return LBT_OK;
}

// Bench / test knob (not in lbt.h): 0 = plain stream-ordered launches, 1 (default) = programmatic dependent launch.
extern "C" int lbt_set_pdl(int on) {
  lbt::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return LBT_OK;
}
