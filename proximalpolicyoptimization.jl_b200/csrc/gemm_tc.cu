// gemm_tc.cu — tcgen05 engines (placeholder until the kernels land: every entry fails loudly).
#include "gemm_tc.cuh"

namespace ppo {

int tc_prepare(ppo_policy*, int) {
    set_error("tensor-core GEMM engines are not built into this library yet");
    return PPO_ERR_STATE;
}
int tc_refresh_weights(ppo_policy*) { set_error("tc engine missing"); return PPO_ERR_STATE; }
int tc_linear_fwd(ppo_policy*, int, const float*, float*, int64_t) { set_error("tc engine missing"); return PPO_ERR_STATE; }
int tc_linear_bwd(ppo_policy*, int, const float*, const float*, float*, float*, float*, int64_t) {
    set_error("tc engine missing");
    return PPO_ERR_STATE;
}
void tc_destroy(ppo_policy*) {}

}  // namespace ppo
