// double instantiation of the env kernels (validation build: compiled with -fmad=false so that every
// product/sum rounds exactly like the float64 numpy reference).
#include "env_kernels.cuh"

namespace mpe {
cudaError_t launch_reset_f64(const EnvStateAny &a, const uint8_t *mask, void *obs, int auto_len, cudaStream_t st) {
  return launch_reset_t<double>(a, mask, obs, auto_len, st);
}
cudaError_t launch_observe_f64(const EnvStateAny &a, void *obs, cudaStream_t st) {
  return launch_observe_t<double>(a, obs, st);
}
cudaError_t launch_step_f64(const EnvStateAny &a, const int32_t *act_u, const int32_t *act_c, const void *comm_vec,
                            void *obs, void *rew, uint8_t *done, int32_t *info_i, void *info_f, cudaStream_t st) {
  return launch_step_t<double>(a, act_u, act_c, comm_vec, obs, rew, done, info_i, info_f, st);
}
cudaError_t launch_set_state_f64(const EnvStateAny &a, const void *pos, const void *vel, const void *lm,
                                 const int32_t *goal, cudaStream_t st) {
  return launch_set_state_t<double>(a, pos, vel, lm, goal, st);
}
cudaError_t launch_get_state_f64(const EnvStateAny &a, void *pos, void *vel, void *lm, int32_t *goal,
                                 cudaStream_t st) {
  return launch_get_state_t<double>(a, pos, vel, lm, goal, st);
}
}  // namespace mpe
