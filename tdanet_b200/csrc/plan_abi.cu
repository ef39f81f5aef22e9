// C-ABI queries about the workspace plan (sizes, latent lengths, named intermediates).
// Compiled into the CUDA library and into the CPU emulation build of the backward pass.
#include "plan.h"

using namespace td;

extern "C" {

int tdanet_workspace_bytes(const tdanet_config_t* cfg, int batch, int n_samples, size_t* bytes) {
  Plan p;
  if (int e = make_plan(cfg, batch, n_samples, p)) return e;
  TD_REQUIRE(bytes != nullptr, "bytes is NULL");
  *bytes = p.bytes;
  return 0;
}

int tdanet_train_workspace_bytes(const tdanet_config_t* cfg, int batch, int n_samples, size_t* bytes) {
  Plan p;
  if (int e = make_plan(cfg, batch, n_samples, p, true)) return e;
  TD_REQUIRE(bytes != nullptr, "bytes is NULL");
  *bytes = p.bytes;
  return 0;
}

int tdanet_latent_lengths(const tdanet_config_t* cfg, int n_samples, int32_t* lengths, int32_t* padded_len, int32_t* rest) {
  Plan p;
  if (int e = make_plan(cfg, 1, n_samples, p)) return e;
  if (lengths) for (int k = 0; k < cfg->depth; ++k) lengths[k] = p.L[k];
  if (padded_len) *padded_len = p.Tp;
  if (rest) *rest = p.rest;
  return 0;
}

static int find_named(const Plan& p, const char* name, int block, size_t* byte_offset, int64_t dims[3], int32_t* elem_bytes) {
  TD_REQUIRE(name != nullptr, "name is NULL");
  TD_REQUIRE(block >= 0 && block < p.n_blk, "block %d outside [0, %d)", block, p.n_blk);
  for (auto it = p.named.rbegin(); it != p.named.rend(); ++it)
    if (it->name == name) {
      if (byte_offset) *byte_offset = p.blk_off(it->off, block);
      if (dims) { dims[0] = it->dims[0]; dims[1] = it->dims[1]; dims[2] = it->dims[2]; }
      if (elem_bytes) *elem_bytes = it->esize;
      return 0;
    }
  return fail(TDANET_EINVAL, "unknown workspace tensor '%s'", name);
}

int tdanet_workspace_tensor(const tdanet_config_t* cfg, int batch, int n_samples, const char* name,
                            size_t* byte_offset, int64_t dims[3]) {
  Plan p;
  if (int e = make_plan(cfg, batch, n_samples, p)) return e;
  return find_named(p, name, 0, byte_offset, dims, nullptr);
}

int tdanet_train_workspace_tensor(const tdanet_config_t* cfg, int batch, int n_samples, const char* name, int block,
                                  size_t* byte_offset, int64_t dims[3], int32_t* elem_bytes) {
  Plan p;
  if (int e = make_plan(cfg, batch, n_samples, p, true)) return e;
  return find_named(p, name, block, byte_offset, dims, elem_bytes);
}

#ifdef TD_EMU
// Test hook of the CPU emulation build (tests/test_deterministic_emu.py): the fixed-point accumulation of the
// deterministic-statistics mode (common.cuh: det_add / det_value) applied to v[order[0]], v[order[1]], ... into the
// pair of a double slot and of a float slot; returns what det_finalize_kernel would write.
int tdanet_emu_det_sum(const double* v, const int32_t* order, int n, double* as_double, float* as_float) {
  alignas(16) char slots[16] = {};          // [0, 8): a double slot, [8, 12): a float slot
  alignas(16) unsigned long long shadow[8] = {};
  const DetRef d{slots, reinterpret_cast<char*>(shadow)};
  for (int i = 0; i < n; ++i) {
    det_add(d, slots, v[order[i]]);
    det_add(d, slots + 8, (double)(float)v[order[i]]);
  }
  if (as_double) *as_double = det_value(shadow);
  if (as_float) *as_float = (float)det_value(shadow + 4);
  return 0;
}
#endif

}  // extern "C"
