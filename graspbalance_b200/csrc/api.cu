// api.cu -- library-level entry points of libgbops: version, error strings, tuning knobs, launch counter.
#include <string.h>

#include "common.cuh"

#include <mutex>
#include <unordered_map>

namespace gb {
std::atomic<unsigned long long> g_launch_count{0};
Tuning g_tuning;

int num_sms() {
  static std::atomic<int> sms[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return kNumSMsB200;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMsB200;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int raise_smem_limit(const void *kernel, size_t needed) {
  static std::mutex mu;
  static std::unordered_map<uintptr_t, size_t> limits;  // (kernel, device) -> dynamic shared-memory limit in force
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const uintptr_t key = reinterpret_cast<uintptr_t>(kernel) * (uintptr_t)kMaxDevices + (uintptr_t)(dev % kMaxDevices);
  std::lock_guard<std::mutex> lock(mu);
  auto it = limits.find(key);
  if (it == limits.end()) {
    cudaFuncAttributes fa;
    int optin = 0;
    if ((e = cudaFuncGetAttributes(&fa, kernel)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return (int)e;
    const size_t lim = (size_t)optin > fa.sharedSizeBytes ? (size_t)optin - fa.sharedSizeBytes : 0;
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim)) != cudaSuccess) return (int)e;
    it = limits.emplace(key, lim).first;
  }
  return needed <= it->second ? 0 : (int)cudaErrorInvalidValue;
}
}  // namespace gb

extern "C" int gb_abi_version(void) { return GBOPS_ABI_VERSION; }

extern "C" const char *gb_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

extern "C" uint64_t gb_launch_count(void) { return (uint64_t)gb::g_launch_count.load(std::memory_order_relaxed); }

static std::atomic<int> *tuning_slot(const char *key) {
  if (!key) return nullptr;
  if (!strcmp(key, "fps_cluster")) return &gb::g_tuning.fps_cluster;
  if (!strcmp(key, "fps_threads")) return &gb::g_tuning.fps_threads;
  if (!strcmp(key, "fps_direct")) return &gb::g_tuning.fps_direct;
  if (!strcmp(key, "fps_defer")) return &gb::g_tuning.fps_defer;
  if (!strcmp(key, "group_split")) return &gb::g_tuning.group_split;
  if (!strcmp(key, "group_mode")) return &gb::g_tuning.group_mode;
  if (!strcmp(key, "group_target_kb")) return &gb::g_tuning.group_target_kb;
  if (!strcmp(key, "group_ch")) return &gb::g_tuning.group_ch;
  if (!strcmp(key, "interp_mode")) return &gb::g_tuning.interp_mode;
  if (!strcmp(key, "query_qpw")) return &gb::g_tuning.query_qpw;
  if (!strcmp(key, "scatter_cc")) return &gb::g_tuning.scatter_cc;
  if (!strcmp(key, "scatter_nt")) return &gb::g_tuning.scatter_nt;
  if (!strcmp(key, "scatter_mode")) return &gb::g_tuning.scatter_mode;
  if (!strcmp(key, "priv_vl")) return &gb::g_tuning.priv_vl;
  if (!strcmp(key, "priv_dry")) return &gb::g_tuning.priv_dry;
  if (!strcmp(key, "priv_split")) return &gb::g_tuning.priv_split;
  if (!strcmp(key, "priv_rows")) return &gb::g_tuning.priv_rows;
  if (!strcmp(key, "priv_cw")) return &gb::g_tuning.priv_cw;
  if (!strcmp(key, "query_mode")) return &gb::g_tuning.query_mode;
  if (!strcmp(key, "grid_cell_pct")) return &gb::g_tuning.grid_cell_pct;
  return nullptr;
}

extern "C" int gb_set_tuning(const char *key, int value) {
  std::atomic<int> *p = tuning_slot(key);
  if (!p) return (int)cudaErrorInvalidValue;
  p->store(value, std::memory_order_relaxed);
  return 0;
}

extern "C" int gb_get_tuning(const char *key, int *value) {
  std::atomic<int> *p = tuning_slot(key);
  if (!p || !value) return (int)cudaErrorInvalidValue;
  *value = p->load(std::memory_order_relaxed);
  return 0;
}
