// placeholder until the tcgen05 window-attention kernel lands
#include "common.cuh"
namespace tmae {
int attn_tc_fwd(const void*, const void*, const void*, void*, float*, const tmae_layer_tables*, const float*, float, int64_t, int64_t, int, int, int, int, int, cudaStream_t) {
  set_error("attn_tc_fwd: not built");
  return TMAE_ERR_UNSUPPORTED;
}
int attn_tc_bwd(const void*, const void*, const void*, const void*, const void*, const float*, const float*, const float*, void*, void*, void*, float*,
                const tmae_layer_tables*, const float*, float, int64_t, int64_t, int, int, int, int, int, cudaStream_t) {
  set_error("attn_tc_bwd: not built");
  return TMAE_ERR_UNSUPPORTED;
}
bool attn_tc_available() { return false; }
}  // namespace tmae
