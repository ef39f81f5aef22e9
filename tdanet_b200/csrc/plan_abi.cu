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

}  // extern "C"
