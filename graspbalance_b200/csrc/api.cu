// api.cu -- library-level entry points of libgbops: version, error strings, tuning knobs, launch counter.
#include <string.h>

#include "common.cuh"

namespace gb {
unsigned long long g_launch_count = 0;
Tuning g_tuning;
}  // namespace gb

extern "C" int gb_abi_version(void) { return GBOPS_ABI_VERSION; }

extern "C" const char *gb_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

extern "C" uint64_t gb_launch_count(void) { return (uint64_t)gb::g_launch_count; }

static int *tuning_slot(const char *key) {
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
  if (!strcmp(key, "query_mode")) return &gb::g_tuning.query_mode;
  if (!strcmp(key, "grid_cell_pct")) return &gb::g_tuning.grid_cell_pct;
  return nullptr;
}

extern "C" int gb_set_tuning(const char *key, int value) {
  int *p = tuning_slot(key);
  if (!p) return (int)cudaErrorInvalidValue;
  *p = value;
  return 0;
}

extern "C" int gb_get_tuning(const char *key, int *value) {
  int *p = tuning_slot(key);
  if (!p || !value) return (int)cudaErrorInvalidValue;
  *value = *p;
  return 0;
}
