// Device-resident replay ring (SURVEY section 8f "next" row 1): rls/replay_buffer.py:9-91 (ReplayBuffer.add /
// make_index / sample_index / _encode_sample) and the transition tuple built at experiments/run.py:46,52
// (obs_n, action_n_env, rew_shared = sum(rew_n), new_obs_n, float(done)), for B env instances at a time.
// Transitions stay in HBM in the caller-facing layouts; both kernels are pure streaming copies/gathers
// (16 B vector accesses, one warp per transition row).
#include "replay_launch.h"

#include "common.cuh"

namespace mpe {

// Both kernels are element-parallel over the flattened [rows][R] observation arrays (8 B vectors when R is even), so
// every lane moves data even though one transition is only R * 4 = 120 B (simple_spread N = 3); the per-row
// extras (action indices, shared reward, done flag) are handled by the first threads of each row.

// append B transitions at ring slots (head + b) % capacity
template <int V>  // V = floats per vector (2 or 1)
__global__ void __launch_bounds__(256) k_replay_add(ReplayDev r, int64_t head, int64_t B, const float *__restrict__ obs,
                                                    const int32_t *__restrict__ act_u, const int32_t *__restrict__ act_c,
                                                    const float *__restrict__ rew, const float *__restrict__ obs_next,
                                                    const float *__restrict__ done) {
  const int R = r.N * r.D, RV = R / V;
  const int64_t total = B * RV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / RV;
    const int c = (int)(i - b * RV);
    const int64_t slot = (head + b) % r.capacity;
    if (V == 2) {
      reinterpret_cast<float2 *>(r.obs + slot * R)[c] = reinterpret_cast<const float2 *>(obs + b * R)[c];
      reinterpret_cast<float2 *>(r.obs_next + slot * R)[c] = reinterpret_cast<const float2 *>(obs_next + b * R)[c];
    } else {
      r.obs[slot * R + c] = obs[b * R + c];
      r.obs_next[slot * R + c] = obs_next[b * R + c];
    }
    if (c < r.N) {
      r.act_u[slot * r.N + c] = (int8_t)act_u[b * r.N + c];
      r.act_c[slot * r.N + c] = act_c != nullptr ? (int8_t)act_c[b * r.N + c] : (int8_t)0;
    }
    if (c == 0) {
      float s = 0.0f;  // rew_shared = np.sum(rew_n) (experiments/run.py:46), agent order
      for (int n = 0; n < r.N; ++n) s += rew[b * r.N + n];
      r.rew[slot] = s;
      r.done[slot] = done != nullptr ? done[b] : 0.0f;
    }
  }
}

// uniform indices with replacement: random.randint(0, len - 1) per sample (rls/replay_buffer.py:51-52)
__global__ void __launch_bounds__(256) k_replay_make_index(int64_t size, int64_t batch, uint64_t seed, uint64_t counter,
                                                           int64_t *__restrict__ idx) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < batch; j += (int64_t)gridDim.x * blockDim.x) {
    const uint4 rr = philox_raw(seed, (uint64_t)j, (uint32_t)counter, 4u, (uint32_t)(counter >> 32) & 0xFFFFu);
    idx[j] = (int64_t)__umul64hi(((uint64_t)rr.x << 32) | rr.y, (uint64_t)size);
  }
}

// gather `batch` transitions by index
template <int V>
__global__ void __launch_bounds__(256) k_replay_gather(ReplayDev r, int64_t size, int64_t batch, const int64_t *__restrict__ idx,
                                                       float *__restrict__ obs, float *__restrict__ act_onehot,
                                                       float *__restrict__ rew, float *__restrict__ obs_next,
                                                       float *__restrict__ done) {
  const int R = r.N * r.D, RV = R / V, A = r.A0 + r.A1;
  const int64_t total = batch * RV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / RV;
    const int c = (int)(i - j * RV);
    // list indexing of ReplayBuffer._storage[i] (rls/replay_buffer.py:42): negative indices count from the end;
    // anything still outside [0, size) is clamped so that a stale index can never read outside the ring
    int64_t slot = idx[j];
    if (slot < 0) slot += size;
    slot = slot < 0 ? 0 : (slot >= size ? size - 1 : slot);
    if (V == 2) {
      if (obs != nullptr) reinterpret_cast<float2 *>(obs + j * R)[c] = reinterpret_cast<const float2 *>(r.obs + slot * R)[c];
      if (obs_next != nullptr)
        reinterpret_cast<float2 *>(obs_next + j * R)[c] = reinterpret_cast<const float2 *>(r.obs_next + slot * R)[c];
    } else {
      if (obs != nullptr) obs[j * R + c] = r.obs[slot * R + c];
      if (obs_next != nullptr) obs_next[j * R + c] = r.obs_next[slot * R + c];
    }
    if (act_onehot != nullptr)
      for (int e = c; e < r.N * A; e += RV) {
        const int n = e / A, a = e - n * A;
        const int u = r.act_u[slot * r.N + n], cc = r.act_c[slot * r.N + n];
        act_onehot[j * r.N * A + e] = (a < r.A0 ? a == u : a - r.A0 == cc) ? 1.0f : 0.0f;
      }
    if (c == 0) {
      if (rew != nullptr) rew[j] = r.rew[slot];
      if (done != nullptr) done[j] = r.done[slot];
    }
  }
}

static unsigned grid_for(int64_t total) {
  const int64_t blocks = (total + 255) / 256;
  return (unsigned)(blocks < 148 * 32 ? (blocks > 0 ? blocks : 1) : 148 * 32);
}

cudaError_t launch_replay_add(const ReplayDev &r, int64_t head, int64_t B, const float *obs, const int32_t *act_u,
                              const int32_t *act_c, const float *rew, const float *obs_next, const float *done,
                              cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const int R = r.N * r.D;
  const bool v2 = (R % 2 == 0) && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(obs_next)) & 7) == 0;
  if (v2)
    k_replay_add<2><<<grid_for(B * (R / 2)), 256, 0, st>>>(r, head, B, obs, act_u, act_c, rew, obs_next, done);
  else
    k_replay_add<1><<<grid_for(B * R), 256, 0, st>>>(r, head, B, obs, act_u, act_c, rew, obs_next, done);
  return cudaGetLastError();
}

cudaError_t launch_replay_make_index(int64_t size, int64_t batch, uint64_t seed, uint64_t counter, int64_t *idx,
                                     cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  k_replay_make_index<<<grid_for(batch), 256, 0, st>>>(size, batch, seed, counter, idx);
  return cudaGetLastError();
}

cudaError_t launch_replay_gather(const ReplayDev &r, int64_t size, int64_t batch, const int64_t *idx, float *obs, float *act_onehot,
                                 float *rew, float *obs_next, float *done, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  const int R = r.N * r.D;
  const bool v2 = (R % 2 == 0) && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(obs_next)) & 7) == 0;
  if (v2)
    k_replay_gather<2><<<grid_for(batch * (R / 2)), 256, 0, st>>>(r, size, batch, idx, obs, act_onehot, rew, obs_next, done);
  else
    k_replay_gather<1><<<grid_for(batch * R), 256, 0, st>>>(r, size, batch, idx, obs, act_onehot, rew, obs_next, done);
  return cudaGetLastError();
}

}  // namespace mpe
