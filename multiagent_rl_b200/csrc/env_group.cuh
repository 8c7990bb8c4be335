// simple_spread step for larger teams (N = 6, 9, 12): G lanes cooperate on one env, each lane owns A = N / G
// agents and A landmarks in registers; positions travel between the lanes of an env by warp shuffle.
//
// Why: the thread-per-env kernel needs every entity of the env in one thread's registers (255 registers and
// spills at N >= 9, 8 warps/SM) and walks N(N-1)/2 pairs serially.  Here a lane keeps only its OWN agents (no atomics,
// G x the parallelism, 56-64 registers).
//   fp64 (validation build): every pair is evaluated by both owners and a lane adds the contributions in the order of
//     the other agent's index, which is upstream's (a, b) lexicographic pair order seen from agent i - the force on
//     agent i from the pair {i, j} is 100 (p_i - p_j) / dist * pen for either ordering, because upstream's f_b = -f
//     with delta = p_a - p_b negates exactly: same bits as the thread-per-env kernel;
//   fp32: most pairs ONCE (grp_physics: own-block pairs once, one partner block per block distance with the negated
//     forces handed over by shuffle): half the contact evaluations at N = 9, a different accumulation order (fp32
//     tolerance against the oracle; k_step_grp and the fused large-team rollout share the code and stay bit-identical);
//   the landmark terms of the reward are summed in landmark order by every lane.
// Reference rows: World.apply_environment_force / get_collision_force / integrate_state, simple_spread
// Scenario.reward / benchmark_data (multiagent package, called from experiments/run.py:44) and
// experiments/scenarios.py:6-20 (observation); make_world(num_agents=n) at experiments/scenarios.py:170.
#pragma once
#include "env_core.cuh"

namespace mpe {

constexpr int kGroupStepThreads = 128;  // block size of k_step_grp / k_reset_grp (== kStepThreads of env_kernels.cuh)

// Team size -> kernel family.  Up to 5 agents: one thread per env (every entity in registers).  6 and more: G lanes
// per env, each owning N / G agents and landmarks; G is the divisor of N that keeps a lane at <= 3 agents with the
// most envs per warp (primes get one agent per lane).
__host__ __device__ constexpr int group_lanes(int N) {
  return N <= 5 ? 0 : N == 6 ? 2 : N == 7 ? 7 : N == 8 ? 4 : N == 9 ? 3 : N == 10 ? 5 : N == 11 ? 11 : N == 12 ? 4 : -1;
}

template <typename T, int N, int G>
struct GroupLayout {
  static constexpr int A = N / G;       // agents (and landmarks) per lane
  static constexpr int EPW = 32 / G;    // envs per warp
  static constexpr int LANES = EPW * G; // active lanes
  static constexpr int D = 4 + 2 * N;
  static constexpr int R = N * D;
  // Staging rows.  The fp32 build writes them with the widest vector the row length allows (float4 when D % 4 == 0,
  // else float2), unpadded, so that the warp's span leaves with ONE TMA bulk store.  Lane (env el, part q) starts at
  // el * R + q * A * D floats; the lanes of one shared-memory wavefront (8 lanes for 16 B stores, 16 for 8 B) then hit
  //   N = 9  (G 3, D 22, R 198):  8 B slot  99 el + 33 q = 3 el + q (mod 16): all distinct
  //   N = 12 (G 4, D 28, R 336): 16 B slot  84 el + 21 q = 4 el + 5 q (mod 8): all distinct
  //   N = 6  (G 2, D 16, R  96): 16 B slot  24 el + 12 q = 4 q (mod 8): 4-way conflicts - still faster than padding
  //          the rows apart and issuing one bulk store per env (measured 0.850 vs 0.825 of HBM: 16 bulk stores per
  //          warp run into the TMA issue rate).
  // The fp64 validation build keeps scalar stores, a 16 B pad and one bulk store per env.
  static constexpr bool kF32 = sizeof(T) == 4;
  static constexpr int kVec = kF32 ? (D % 4 == 0 ? 4 : 2) : 1;
  static constexpr int kPad = kF32 ? 0 : ((R * (int)sizeof(T)) % 16 == 0 ? 16 / (int)sizeof(T) : 0);
  static constexpr bool kPerEnv = kPad != 0;
  static constexpr int RS = R + kPad;
  // rewards are NOT staged (3 plain stores per lane): the tile is the occupancy limit of these kernels (N = 9: 8,320 ->
  // 7,936 B per warp = 7 instead of 6 resident CTAs per SM)
  static constexpr int kWarpBytes = (EPW * RS * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr int kBlockBytes = kWarpBytes * (kGroupStepThreads / 32);
  static_assert(N % G == 0, "agents must split evenly over the lanes of an env");
};

// lane bookkeeping of the lanes-per-env kernels
template <typename T, int N, int G>
struct GroupLanes {
  int warp, lane, el, q, base_lane;
  int64_t b0, b;
  bool lane_ok, active, full;
};

// Observation rows of the lane's own agents (+ optional rewards) -> caller-facing obs[b][N][D] / rew[b][N]; a warp's
// envs are one contiguous span of both, staged in shared memory and handed to the TMA engine.  Returns whether this
// lane issued bulk stores (it must then call bulk_wait_read_all() before the kernel ends).
template <typename T, int N, int G>
__device__ __forceinline__ bool grp_emit(const GroupLanes<T, N, G> &g, const T (&px)[N / G], const T (&py)[N / G],
                                         const T (&vx)[N / G], const T (&vy)[N / G], const T (&lx)[N / G],
                                         const T (&ly)[N / G], const T *r, T *obs, T *rew, unsigned char *smem) {
  using GL = GroupLayout<T, N, G>;
  constexpr int A = GL::A, EPW = GL::EPW, D = GL::D, R = GL::R, L = N;
  const unsigned FULL = 0xffffffffu;
  const int warp = g.warp, lane = g.lane, el = g.el, q = g.q, base_lane = g.base_lane;
  const int64_t b0 = g.b0, b = g.b;
  const bool lane_ok = g.lane_ok, active = g.active, full = g.full;
  constexpr int RS = GL::RS;
  T *st_obs = reinterpret_cast<T *>(smem + warp * GL::kWarpBytes);
  T *st_rew = st_obs + EPW * RS;
  // smem == nullptr: no staging buffer (the fused rollout's env phase) - rows go straight to global memory
  const bool obs_tma = smem != nullptr && full && obs != nullptr && (reinterpret_cast<uintptr_t>(obs + b0 * R) & 15) == 0 &&
                       (GL::kPerEnv || (EPW * R * sizeof(T)) % 16 == 0);
  const bool rew_tma = false;  // see GroupLayout::kWarpBytes
  if (obs != nullptr && obs_tma && GL::kVec > 1) {
    // shared-memory staging with vector stores (the pointer is derived from the shared array only, so these are STS)
    float *rowbase = reinterpret_cast<float *>(smem + warp * GL::kWarpBytes) + el * RS + q * A * D;
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < A; ++k) {
        if (GL::kVec == 4) {
          *reinterpret_cast<float4 *>(rowbase + k * D) = make_float4((float)vx[k], (float)vy[k], (float)px[k], (float)py[k]);
        } else {
          *reinterpret_cast<float2 *>(rowbase + k * D) = make_float2((float)vx[k], (float)vy[k]);
          *reinterpret_cast<float2 *>(rowbase + k * D + 2) = make_float2((float)px[k], (float)py[k]);
        }
      }
    }
    if (GL::kVec == 4) {
#pragma unroll
      for (int l = 0; l < L; l += 2) {
        const T ax = __shfl_sync(FULL, lx[l % A], base_lane + l / A), ay = __shfl_sync(FULL, ly[l % A], base_lane + l / A);
        const T bx = __shfl_sync(FULL, lx[(l + 1) % A], base_lane + (l + 1) / A);
        const T by = __shfl_sync(FULL, ly[(l + 1) % A], base_lane + (l + 1) / A);
        if (lane_ok) {
#pragma unroll
          for (int k = 0; k < A; ++k)
            *reinterpret_cast<float4 *>(rowbase + k * D + 4 + 2 * l) =
                make_float4((float)(ax - px[k]), (float)(ay - py[k]), (float)(bx - px[k]), (float)(by - py[k]));
        }
      }
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const T llx = __shfl_sync(FULL, lx[l % A], base_lane + l / A);
        const T lly = __shfl_sync(FULL, ly[l % A], base_lane + l / A);
        if (lane_ok) {
#pragma unroll
          for (int k = 0; k < A; ++k)
            *reinterpret_cast<float2 *>(rowbase + k * D + 4 + 2 * l) = make_float2((float)(llx - px[k]), (float)(lly - py[k]));
        }
      }
    }
  } else if (obs != nullptr) {
    T *rowbase = obs_tma ? st_obs + el * RS : obs + b * R;
    const bool wr = obs_tma ? lane_ok : active;
    if (wr) {
#pragma unroll
      for (int k = 0; k < A; ++k) {
        T *row = rowbase + (q * A + k) * D;
        row[0] = vx[k]; row[1] = vy[k]; row[2] = px[k]; row[3] = py[k];
      }
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const T llx = __shfl_sync(FULL, lx[l % A], base_lane + l / A);
      const T lly = __shfl_sync(FULL, ly[l % A], base_lane + l / A);
      if (wr) {
#pragma unroll
        for (int k = 0; k < A; ++k) {
          T *row = rowbase + (q * A + k) * D;
          row[4 + 2 * l] = llx - px[k];
          row[5 + 2 * l] = lly - py[k];
        }
      }
    }
  }
  if (rew != nullptr && (rew_tma ? lane_ok : active)) {
    T *rr = rew_tma ? st_rew + el * N : rew + b * N;
#pragma unroll
    for (int k = 0; k < A; ++k) rr[q * A + k] = r[k];
  }
  bool issued = false;
  if (obs_tma || rew_tma) {
    fence_proxy_async_smem();
    __syncwarp();
    if (GL::kPerEnv) {
      if (obs_tma && lane_ok && q == 0) {  // one bulk store per env (rows are padded apart in smem)
        bulk_store(obs + b * R, st_obs + el * RS, R * sizeof(T));
        issued = true;
      }
    } else if (obs_tma && lane == 0) {
      bulk_store(obs + b0 * R, st_obs, EPW * R * sizeof(T));
      issued = true;
    }
    if (rew_tma && lane == 0) {
      bulk_store(rew + b0 * N, st_rew, EPW * N * sizeof(T));
      issued = true;
    }
    if (issued) bulk_commit();
  }
  return issued;
}

// ---- pieces of the step shared by k_step_grp and the fused large-team rollout (tc_kernels.cu): every function below
// ---- must be called by all 32 lanes of the warp (warp shuffles), `q` = lane within the env, `base_lane` = its first lane

// the lane's agents and landmarks of env b (SoA state); inactive lanes keep zeros
template <typename T, int N, int G>
__device__ __forceinline__ void grp_load(const EnvState<T> &s, int64_t b, int q, bool active, T (&px)[N / G], T (&py)[N / G],
                                         T (&vx)[N / G], T (&vy)[N / G], T (&lx)[N / G], T (&ly)[N / G]) {
  constexpr int A = N / G;
#pragma unroll
  for (int k = 0; k < A; ++k) px[k] = py[k] = vx[k] = vy[k] = lx[k] = ly[k] = (T)0;
  if (active) {
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const int i = q * A + k;
      const Vec4<T> v = ld4(s.pv + ((int64_t)i * s.B + b) * 4);
      px[k] = v.x; py[k] = v.y; vx[k] = v.z; vy[k] = v.w;
      const Vec2<T> l = ld2(s.lm + ((int64_t)i * s.B + b) * 2);
      lx[k] = l.x; ly[k] = l.y;
    }
  }
}

// _set_action + apply_environment_force on the lane's own agents + integrate_state
template <typename T, int N, int G>
__device__ __forceinline__ void grp_physics(const EnvState<T> &s, const int (&au)[N / G], int q, int base_lane,
                                            T (&px)[N / G], T (&py)[N / G], T (&vx)[N / G], T (&vy)[N / G]) {
  constexpr int A = N / G;
  const unsigned FULL = 0xffffffffu;
  const T sens = s.accel >= (T)0 ? s.accel : (T)5.0;
  T fx[A], fy[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const int a = au[k];
    fx[k] = ((T)0 + ((a == 1 ? (T)1 : (T)0) - (a == 2 ? (T)1 : (T)0))) * sens;
    fy[k] = ((T)0 + ((a == 3 ? (T)1 : (T)0) - (a == 4 ? (T)1 : (T)0))) * sens;
  }
  const T dist_min = (T)0.15 + (T)0.15;
  if constexpr (std::is_same<T, float>::value && (N / G) > 1) {
    // fp32: most pairs ONCE (the force of a pair negates exactly between its two agents).  Pairs inside the lane's own
    // block are evaluated once and applied to both agents; for every block distance d < G / 2 lane q evaluates the
    // A x A pairs between its block and block q + d and hands the (negated) forces to that block's lane, receiving
    // block q - d's in the same shuffles; only the block at distance G / 2 (even G) is evaluated from both sides.
    // N = 9: 12 contact evaluations per lane instead of 24, N = 12: 21 instead of 33.  The accumulation order differs
    // from upstream's (fp32 tolerance); the fp64 build below keeps it.  Used by k_step_grp AND the fused large-team
    // rollout, which therefore stay bit-identical to each other.
#pragma unroll
    for (int k = 0; k < A; ++k)
#pragma unroll
      for (int m = k + 1; m < A; ++m) {
        const T dx = px[k] - px[m], dy = py[k] - py[m];
        T gx, gy;
        contact_force<T>(dx, dy, sq2<T>(dx, dy), dist_min, gx, gy);
        fx[k] += gx; fy[k] += gy;
        fx[m] -= gx; fy[m] -= gy;
      }
#pragma unroll
    for (int d = 1; 2 * d <= G; ++d) {
      const int pb = q + d >= G ? q + d - G : q + d;   // block whose agents this lane evaluates against
      const int rb = q - d < 0 ? q - d + G : q - d;    // block whose lane evaluated against this lane's agents
      T ox[A], oy[A];
#pragma unroll
      for (int m = 0; m < A; ++m) { ox[m] = __shfl_sync(FULL, px[m], base_lane + pb); oy[m] = __shfl_sync(FULL, py[m], base_lane + pb); }
      T gx[A][A], gy[A][A];
#pragma unroll
      for (int k = 0; k < A; ++k)
#pragma unroll
        for (int m = 0; m < A; ++m) {
          const T dx = px[k] - ox[m], dy = py[k] - oy[m];
          contact_force<T>(dx, dy, sq2<T>(dx, dy), dist_min, gx[k][m], gy[k][m]);
          fx[k] += gx[k][m]; fy[k] += gy[k][m];
        }
      if (2 * d < G) {  // compile time: the opposite block of an even G is evaluated from both sides instead
#pragma unroll
        for (int k = 0; k < A; ++k)
#pragma unroll
          for (int m = 0; m < A; ++m) {  // lane rb's force on ITS agent k from MY agent m: mine is the negative
            fx[m] -= __shfl_sync(FULL, gx[k][m], base_lane + rb);
            fy[m] -= __shfl_sync(FULL, gy[k][m], base_lane + rb);
          }
      }
    }
  } else {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const T pjx = __shfl_sync(FULL, px[j % A], base_lane + j / A);
    const T pjy = __shfl_sync(FULL, py[j % A], base_lane + j / A);
#pragma unroll
    for (int k = 0; k < A; ++k) {
      if (q * A + k == j) continue;
      const T dx = px[k] - pjx, dy = py[k] - pjy;
      const T d2 = sq2<T>(dx, dy);
      if (std::is_same<T, float>::value || !(d2 >= s.t2_cut)) {
        T gx, gy;
        contact_force<T>(dx, dy, d2, dist_min, gx, gy);
        fx[k] = gx + fx[k];
        fy[k] = gy + fy[k];
      }
    }
  }
  }
#pragma unroll
  for (int k = 0; k < A; ++k) integrate_agent<T>(px[k], py[k], vx[k], vy[k], fx[k], fy[k], s.max_speed);
}

// reward: nearest agent of the lane's own landmarks, collisions of the lane's own agents; r[k] per own agent,
// occ / md (benchmark_data) identical on every lane of the env
template <typename T, int N, int G>
__device__ __forceinline__ void grp_reward(const EnvState<T> &s, int base_lane, const T (&px)[N / G], const T (&py)[N / G],
                                           const T (&lx)[N / G], const T (&ly)[N / G], T (&r)[N / G], int (&coll)[N / G],
                                           int &occ, T &md) {
  constexpr int A = N / G, L = N;
  const unsigned FULL = 0xffffffffu;
  T m2[A];
#pragma unroll
  for (int k = 0; k < A; ++k) coll[k] = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const T pjx = __shfl_sync(FULL, px[j % A], base_lane + j / A);
    const T pjy = __shfl_sync(FULL, py[j % A], base_lane + j / A);
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const T dx = pjx - lx[k], dy = pjy - ly[k];
      const T d2 = sq2<T>(dx, dy);
      m2[k] = (j == 0) ? d2 : (d2 < m2[k] ? d2 : m2[k]);
      const T cx = pjx - px[k], cy = pjy - py[k];  // includes j == own agent: dist 0 < dist_min (NaN: no hit)
      coll[k] += (sq2<T>(cx, cy) < s.t2_coll) ? 1 : 0;
    }
  }
  T m[A];
#pragma unroll
  for (int k = 0; k < A; ++k) m[k] = sqrt(m2[k]);
  T base = (T)0;
  md = (T)0;
  occ = 0;
#pragma unroll
  for (int l = 0; l < L; ++l) {  // landmark order, like upstream's loop
    const T ml = __shfl_sync(FULL, m[l % A], base_lane + l / A);
    const T ml2 = __shfl_sync(FULL, m2[l % A], base_lane + l / A);
    base -= ml;
    md += ml;
    occ += (ml2 < s.t2_occ) ? 1 : 0;
  }
#pragma unroll
  for (int k = 0; k < A; ++k) {
    T rr = base;
    for (int c = 0; c < coll[k]; ++c) rr -= (T)1;
    r[k] = rr;
  }
}

// team reward = sum over the lanes of the env, added in lane order (convergent shuffles); same value on every lane
template <typename T, int N, int G>
__device__ __forceinline__ T grp_team_sum(const T (&r)[N / G], int base_lane) {
  constexpr int A = N / G;
  T sum = (T)0;
#pragma unroll
  for (int k = 0; k < A; ++k) sum += r[k];
  T tot = (T)0;
#pragma unroll
  for (int g = 0; g < G; ++g) tot += __shfl_sync(0xffffffffu, sum, base_lane + g);
  return tot;
}

// Scenario.reset_world for the lane's entities of env b, episode ep (same draws as Env::reset: agents, then landmarks)
template <typename T, int N, int G>
__device__ __forceinline__ void grp_reset_draw(const EnvState<T> &s, int64_t b, uint32_t ep, int q, T (&px)[N / G],
                                               T (&py)[N / G], T (&vx)[N / G], T (&vy)[N / G], T (&lx)[N / G], T (&ly)[N / G]) {
  constexpr int A = N / G;
  const uint64_t gid = (uint64_t)(s.gid0 + b);
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const int ia = q * A + k, il = N + q * A + k;  // entity indices of the lane's k-th agent / landmark
    const uint4 ra = philox_raw(s.seed, gid, ep, kDomainReset, ia >> 1);
    px[k] = bits_to_pos<T>((ia & 1) ? ra.z : ra.x); py[k] = bits_to_pos<T>((ia & 1) ? ra.w : ra.y);
    const uint4 rl = philox_raw(s.seed, gid, ep, kDomainReset, il >> 1);
    lx[k] = bits_to_pos<T>((il & 1) ? rl.z : rl.x); ly[k] = bits_to_pos<T>((il & 1) ? rl.w : rl.y);
    vx[k] = vy[k] = (T)0;
    st4(s.pv + ((int64_t)ia * s.B + b) * 4, Vec4<T>{px[k], py[k], (T)0, (T)0});
    st2(s.lm + ((int64_t)(q * A + k) * s.B + b) * 2, Vec2<T>{lx[k], ly[k]});
  }
}

template <typename T, int N, int G>
__global__ void __launch_bounds__(kGroupStepThreads)
    k_step_grp(EnvState<T> s, const int32_t *__restrict__ act_u, T *__restrict__ obs, T *__restrict__ rew,
               uint8_t *__restrict__ done, int32_t *__restrict__ info_i, T *__restrict__ info_f) {
  using GL = GroupLayout<T, N, G>;
  constexpr int A = GL::A, EPW = GL::EPW;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int el = lane / G, q = lane - el * G;            // env within the warp, lane within the env
  const int base_lane = el * G;
  const int64_t b0 = ((int64_t)blockIdx.x * (kGroupStepThreads / 32) + warp) * EPW;
  const int64_t b = b0 + el;
  const bool lane_ok = lane < GL::LANES;
  const bool active = lane_ok && b < s.B;
  const bool full = b0 + EPW <= s.B;

  T px[A], py[A], vx[A], vy[A], lx[A], ly[A];
  int au[A];
  grp_load<T, N, G>(s, b, q, active, px, py, vx, vy, lx, ly);
#pragma unroll
  for (int k = 0; k < A; ++k) au[k] = active ? act_u[b * N + q * A + k] : 0;
  grp_physics<T, N, G>(s, au, q, base_lane, px, py, vx, vy);
  if (active) {
#pragma unroll
    for (int k = 0; k < A; ++k) st4(s.pv + ((int64_t)(q * A + k) * s.B + b) * 4, Vec4<T>{px[k], py[k], vx[k], vy[k]});
  }
  T r[A], md;
  int coll[A], occ;
  grp_reward<T, N, G>(s, base_lane, px, py, lx, ly, r, coll, occ, md);
  if (s.track) {
    const T tot = grp_team_sum<T, N, G>(r, base_lane);
    if (active && q == 0) { s.ep_ret[b] += tot; s.tstep[b] += 1; }
  }
  if (active) {
    if (info_i != nullptr) {
#pragma unroll
      for (int k = 0; k < A; ++k) info_i[b * (N + 1) + q * A + k] = coll[k];
      if (q == 0) info_i[b * (N + 1) + N] = occ;
    }
    if (info_f != nullptr && q == 0) info_f[b] = md;
  }

  GroupLanes<T, N, G> gl{warp, lane, el, q, base_lane, b0, b, lane_ok, active, full};
  const bool issued = grp_emit<T, N, G>(gl, px, py, vx, vy, lx, ly, r, obs, rew, smem);
  if (done != nullptr && active) {
#pragma unroll
    for (int k = 0; k < A; ++k) done[b * N + q * A + k] = 0;
  }
  if (issued) bulk_wait_read_all();
}

// env.reset() / scenario.observation for the same lane layout: Philox reset of the masked (or timed-out) envs, then
// the observation rows of every env.  Same draws as Env::reset (entities in upstream's order: agents, landmarks).
template <typename T, int N, int G>
__global__ void __launch_bounds__(kGroupStepThreads)
    k_reset_grp(EnvState<T> s, const uint8_t *__restrict__ mask, T *__restrict__ obs, int auto_len, int do_reset) {
  using GL = GroupLayout<T, N, G>;
  constexpr int A = GL::A, EPW = GL::EPW;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int el = lane / G, q = lane - el * G;
  const int base_lane = el * G;
  const int64_t b0 = ((int64_t)blockIdx.x * (kGroupStepThreads / 32) + warp) * EPW;
  const int64_t b = b0 + el;
  const bool lane_ok = lane < GL::LANES;
  const bool active = lane_ok && b < s.B;
  const bool full = b0 + EPW <= s.B;
  T px[A], py[A], vx[A], vy[A], lx[A], ly[A];
#pragma unroll
  for (int k = 0; k < A; ++k) px[k] = py[k] = vx[k] = vy[k] = lx[k] = ly[k] = (T)0;
  double ret = 0.0, n_ep = 0.0, n_steps = 0.0;
  uint32_t ep_old = 0;
  int t_old = 0;
  bool doit = false;
  if (active) {  // every lane of an env reads the counters before lane 0 of the env rewrites them below
    ep_old = s.episode[b];
    t_old = s.tstep[b];
    doit = do_reset && (mask == nullptr || mask[b] != 0) && (auto_len <= 0 || t_old >= auto_len);
  }
  __syncwarp();
  if (active) {
    if (doit) {
      const uint32_t ep = ep_old + 1u;
      grp_reset_draw<T, N, G>(s, b, ep, q, px, py, vx, vy, lx, ly);
      if (q == 0) {
        if (s.track && t_old > 0) { ret = (double)s.ep_ret[b]; n_ep = 1.0; n_steps = (double)t_old; }
        s.episode[b] = ep;
        s.tstep[b] = 0;
        s.ep_ret[b] = (T)0;
      }
    } else if (obs != nullptr) {
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const Vec4<T> v = ld4(s.pv + ((int64_t)(q * A + k) * s.B + b) * 4);
        px[k] = v.x; py[k] = v.y; vx[k] = v.z; vy[k] = v.w;
        const Vec2<T> l = ld2(s.lm + ((int64_t)(q * A + k) * s.B + b) * 2);
        lx[k] = l.x; ly[k] = l.y;
      }
    }
  }
  if (s.track && do_reset) fold_stats(s.stats, ret, n_ep, n_steps);
  bool issued = false;
  if (obs != nullptr) {
    GroupLanes<T, N, G> gl{warp, lane, el, q, base_lane, b0, b, lane_ok, active, full};
    issued = grp_emit<T, N, G>(gl, px, py, vx, vy, lx, ly, nullptr, obs, nullptr, smem);
  }
  if (issued) bulk_wait_read_all();
}

}  // namespace mpe
