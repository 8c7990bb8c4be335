// Host-side interface of the replay kernels (replay_kernels.cu) used by cabi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpe {

struct ReplayDev {
  float *obs = nullptr, *obs_next = nullptr;  // [capacity][N][D]
  int8_t *act_u = nullptr, *act_c = nullptr;  // [capacity][N] head indices
  float *rew = nullptr, *done = nullptr;      // [capacity] shared reward, done flag
  int64_t capacity = 0;
  int32_t N = 0, D = 0, A0 = 0, A1 = 0;
};

cudaError_t launch_replay_add(const ReplayDev &r, int64_t head, int64_t B, const float *obs, const int32_t *act_u,
                              const int32_t *act_c, const float *rew, const float *obs_next, const float *done,
                              cudaStream_t st);
cudaError_t launch_replay_make_index(int64_t size, int64_t batch, uint64_t seed, uint64_t counter, int64_t *idx,
                                     cudaStream_t st);
cudaError_t launch_replay_gather(const ReplayDev &r, int64_t size, int64_t batch, const int64_t *idx, float *obs, float *act_onehot,
                                 float *rew, float *obs_next, float *done, cudaStream_t st);

}  // namespace mpe
