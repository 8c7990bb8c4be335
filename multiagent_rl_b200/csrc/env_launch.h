// Host-side launch interface between the C ABI (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpe {

// Type-erased view of one shard's persistent device state (see EnvState<T> in env_core.cuh).
struct EnvStateAny {
  void *pv = nullptr, *lm = nullptr, *ep_ret = nullptr, *comm = nullptr;
  int32_t *goal = nullptr;
  uint32_t *episode = nullptr;
  int32_t *tstep = nullptr;
  double *stats = nullptr;
  int64_t B = 0, gid0 = 0;
  uint64_t seed = 0;
  double max_speed = -1.0, accel = -1.0;
  int32_t track = 0;
  int32_t precision = 0, scenario = 0, N = 0, L = 0, D = 0, dimc = 0, act_u = 5, act_c = 0;
  int32_t max_episode_len = 25;
};

bool env_supported(int scenario, int N);

cudaError_t launch_reset(const EnvStateAny &a, const uint8_t *mask, void *obs, int auto_len, cudaStream_t st);
cudaError_t launch_observe(const EnvStateAny &a, void *obs, cudaStream_t st);
cudaError_t launch_step(const EnvStateAny &a, const int32_t *act_u, const int32_t *act_c, const void *comm_vec,
                        void *obs, void *rew, uint8_t *done, int32_t *info_i, void *info_f, cudaStream_t st);
cudaError_t launch_set_state(const EnvStateAny &a, const void *pos, const void *vel, const void *lm,
                             const int32_t *goal, cudaStream_t st);
cudaError_t launch_get_state(const EnvStateAny &a, void *pos, void *vel, void *lm, int32_t *goal, cudaStream_t st);

// per-precision entry points (defined in env_kernels_f32.cu / env_kernels_f64.cu)
#define MPE_DECL_PRECISION(SFX)                                                                              \
  cudaError_t launch_reset_##SFX(const EnvStateAny &, const uint8_t *, void *, int, cudaStream_t);          \
  cudaError_t launch_observe_##SFX(const EnvStateAny &, void *, cudaStream_t);                              \
  cudaError_t launch_step_##SFX(const EnvStateAny &, const int32_t *, const int32_t *, const void *, void *, \
                                void *, uint8_t *, int32_t *, void *, cudaStream_t);                         \
  cudaError_t launch_set_state_##SFX(const EnvStateAny &, const void *, const void *, const void *,          \
                                     const int32_t *, cudaStream_t);                                         \
  cudaError_t launch_get_state_##SFX(const EnvStateAny &, void *, void *, void *, int32_t *, cudaStream_t);
MPE_DECL_PRECISION(f32)
MPE_DECL_PRECISION(f64)
#undef MPE_DECL_PRECISION

}  // namespace mpe
