// Device-resident replay ring (SURVEY section 8f "next" row 1): rls/replay_buffer.py:9-91 (ReplayBuffer.add /
// make_index / sample_index / _encode_sample) and the transition tuple built at experiments/run.py:46,52
// (obs_n, action_n_env, rew_shared = sum(rew_n), new_obs_n, float(done)), for B env instances at a time.
// Transitions stay in HBM in the caller-facing layouts; both kernels are pure streaming copies/gathers
// (16 B vector accesses, one warp per transition row).
#include "replay_launch.h"

#include "common.cuh"

namespace mpe {

// append B transitions at ring slots (head + b) % capacity
__global__ void __launch_bounds__(256) k_replay_add(ReplayDev r, int64_t head, int64_t B, const float *__restrict__ obs,
                                                    const int32_t *__restrict__ act_u, const int32_t *__restrict__ act_c,
                                                    const float *__restrict__ rew, const float *__restrict__ obs_next,
                                                    const float *__restrict__ done) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int R = r.N * r.D;
  for (int64_t b = warp; b < B; b += nwarps) {
    const int64_t slot = (head + b) % r.capacity;
    const float *so = obs + b * R, *sn = obs_next + b * R;
    float *dobs = r.obs + slot * R, *dnext = r.obs_next + slot * R;
    if ((R & 3) == 0 && ((reinterpret_cast<uintptr_t>(so) | reinterpret_cast<uintptr_t>(sn)) & 15) == 0) {
      for (int i = lane; i < R / 4; i += 32) {
        reinterpret_cast<float4 *>(dobs)[i] = reinterpret_cast<const float4 *>(so)[i];
        reinterpret_cast<float4 *>(dnext)[i] = reinterpret_cast<const float4 *>(sn)[i];
      }
    } else {
      for (int i = lane; i < R; i += 32) { dobs[i] = so[i]; dnext[i] = sn[i]; }
    }
    if (lane < r.N) {
      r.act_u[slot * r.N + lane] = (int8_t)act_u[b * r.N + lane];
      r.act_c[slot * r.N + lane] = act_c != nullptr ? (int8_t)act_c[b * r.N + lane] : (int8_t)0;
    }
    if (lane == 0) {
      float s = 0.0f;  // rew_shared = np.sum(rew_n) (experiments/run.py:46), agent order
      for (int i = 0; i < r.N; ++i) s += rew[b * r.N + i];
      r.rew[slot] = s;
      r.done[slot] = done != nullptr ? done[b] : 0.0f;
    }
  }
}

// gather `batch` transitions: idx given, or drawn uniformly with replacement from [0, size) by Philox
__global__ void __launch_bounds__(256) k_replay_sample(ReplayDev r, int64_t size, int64_t batch, const int64_t *__restrict__ idx_in,
                                                       uint64_t seed, uint64_t counter, float *__restrict__ obs,
                                                       float *__restrict__ act_onehot, float *__restrict__ rew,
                                                       float *__restrict__ obs_next, float *__restrict__ done,
                                                       int64_t *__restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int R = r.N * r.D, A = r.A0 + r.A1;
  for (int64_t j = warp; j < batch; j += nwarps) {
    int64_t slot;
    if (idx_in != nullptr) {
      slot = idx_in[j];
    } else {  // random.randint(0, len - 1) with replacement (rls/replay_buffer.py:51-52)
      const uint4 rr = philox_raw(seed, (uint64_t)j, (uint32_t)counter, 4u, (uint32_t)(counter >> 32));
      const uint64_t r64 = ((uint64_t)rr.x << 32) | rr.y;
      slot = (int64_t)__umul64hi(r64, (uint64_t)size);
    }
    if (idx_out != nullptr && lane == 0) idx_out[j] = slot;
    const float *so = r.obs + slot * R, *sn = r.obs_next + slot * R;
    if ((R & 3) == 0 && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(obs_next)) & 15) == 0) {
      for (int i = lane; i < R / 4; i += 32) {
        if (obs != nullptr) reinterpret_cast<float4 *>(obs + j * R)[i] = reinterpret_cast<const float4 *>(so)[i];
        if (obs_next != nullptr) reinterpret_cast<float4 *>(obs_next + j * R)[i] = reinterpret_cast<const float4 *>(sn)[i];
      }
    } else {
      for (int i = lane; i < R; i += 32) {
        if (obs != nullptr) obs[j * R + i] = so[i];
        if (obs_next != nullptr) obs_next[j * R + i] = sn[i];
      }
    }
    if (act_onehot != nullptr)
      for (int i = lane; i < r.N * A; i += 32) {
        const int n = i / A, a = i - n * A;
        const int u = r.act_u[slot * r.N + n], c = r.act_c[slot * r.N + n];
        act_onehot[j * r.N * A + i] = (a < r.A0 ? a == u : a - r.A0 == c) ? 1.0f : 0.0f;
      }
    if (lane == 0) {
      if (rew != nullptr) rew[j] = r.rew[slot];
      if (done != nullptr) done[j] = r.done[slot];
    }
  }
}

cudaError_t launch_replay_add(const ReplayDev &r, int64_t head, int64_t B, const float *obs, const int32_t *act_u,
                              const int32_t *act_c, const float *rew, const float *obs_next, const float *done,
                              cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const int64_t blocks = (B + 7) / 8;  // 8 warps per block, one transition per warp
  k_replay_add<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(r, head, B, obs, act_u, act_c, rew, obs_next, done);
  return cudaGetLastError();
}

cudaError_t launch_replay_sample(const ReplayDev &r, int64_t size, int64_t batch, const int64_t *idx_in, uint64_t seed,
                                 uint64_t counter, float *obs, float *act_onehot, float *rew, float *obs_next, float *done,
                                 int64_t *idx_out, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  const int64_t blocks = (batch + 7) / 8;
  k_replay_sample<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(r, size, batch, idx_in, seed, counter, obs,
                                                                                     act_onehot, rew, obs_next, done, idx_out);
  return cudaGetLastError();
}

}  // namespace mpe
