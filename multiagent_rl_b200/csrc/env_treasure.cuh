// fullobs_collect_treasure: 6 collectors + 2 deposits (all of them agents), 6 treasures.
//
// Reference rows: the scenario is named at main.py:24-25, its observation is the reference's own
// local_obs_collect_treasure (experiments/scenarios.py:95-121: position, velocity, holding one-hot, the six
// treasures nearest first as offset + type one-hot; 30 floats), its post_step hook is handed to MultiAgentEnv at
// experiments/scenarios.py:174-190.  Engine and scenario arithmetic (cached distances, mass ratios in contact
// forces, mass * accel action forces, max_speed, collector / deposit / global rewards, pick-up, respawn, deposit)
// are the MAAC fork's multiagent/core.py and scenarios/fullobs_collect_treasure.py as restated in
// oracle/maac_ref.py (parity unpinned: the fork is not in the reference tree; the four ambiguities are listed there).
//
// One thread per env, every entity in registers.  Per-env integer state is one word (EnvState::goal):
//   bit l        type of treasure l (two types = the two deposits)
//   bit 6 + l    treasure l is alive (a collected treasure sits at (-999, -999) until it respawns one step later)
//   bits 12+2i   what collector i holds: 0 = nothing, 1 + type otherwise
// Observation rows are staged two agents at a time in shared memory (row stride 61: conflict-free scalar stores at
// the rank-dependent offsets of the sorted treasure list) and leave as coalesced 128 B warp stores.
#pragma once
#include "env_core.cuh"

namespace mpe {

constexpr int kTrN = 8, kTrC = 6, kTrL = 6, kTrD = 30, kTrR = kTrN * kTrD;
constexpr int kTrChunk = 2 * kTrD;      // two agents' rows per flush
constexpr int kTrStride = kTrChunk + 1;  // odd shared-memory row stride

template <typename T>
struct TrLayout {
  static constexpr int kObsElems = 32 * kTrStride;
  static constexpr int kWarpBytes = ((kObsElems + 32 * kTrN) * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr int kBlockBytes = kWarpBytes * (kStepThreads / 32);
};

__device__ __forceinline__ int tr_type(uint32_t f, int l) { return (int)((f >> l) & 1u); }
__device__ __forceinline__ bool tr_alive(uint32_t f, int l) { return ((f >> (6 + l)) & 1u) != 0u; }
__device__ __forceinline__ int tr_hold(uint32_t f, int i) { return (int)((f >> (12 + 2 * i)) & 3u) - 1; }  // -1: nothing

template <typename T>
struct TreasureEnv {
  T px[kTrN], py[kTrN], vx[kTrN], vy[kTrN];
  T tx[kTrL], ty[kTrL];
  uint32_t flags;

  __device__ __forceinline__ static constexpr bool collector(int i) { return i < kTrC; }

  __device__ __forceinline__ void load(const EnvState<T> &s, int64_t b) {
#pragma unroll
    for (int i = 0; i < kTrN; ++i) {
      const Vec4<T> v = ld4(s.pv + ((int64_t)i * s.B + b) * 4);
      px[i] = v.x; py[i] = v.y; vx[i] = v.z; vy[i] = v.w;
    }
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      const Vec2<T> v = ld2(s.lm + ((int64_t)l * s.B + b) * 2);
      tx[l] = v.x; ty[l] = v.y;
    }
    flags = (uint32_t)s.goal[b];
  }
  __device__ __forceinline__ void store_agents(const EnvState<T> &s, int64_t b) const {
#pragma unroll
    for (int i = 0; i < kTrN; ++i) st4(s.pv + ((int64_t)i * s.B + b) * 4, Vec4<T>{px[i], py[i], vx[i], vy[i]});
  }
  __device__ __forceinline__ void store_treasures(const EnvState<T> &s, int64_t b, uint32_t moved) const {
#pragma unroll
    for (int l = 0; l < kTrL; ++l)
      if ((moved >> l) & 1u) st2(s.lm + ((int64_t)l * s.B + b) * 2, Vec2<T>{tx[l], ty[l]});
    s.goal[b] = (int32_t)flags;
  }

  // Scenario.reset_world: agents ~ U[-1,1)^2, then per treasure a type and a position ~ 0.95 U[-1,1)^2; nobody holds
  __device__ __forceinline__ void reset(uint64_t seed, uint64_t gid, uint32_t episode) {
    constexpr int E = kTrN + kTrL;
#pragma unroll
    for (int j = 0; j < E / 2; ++j) {
      const uint4 r = philox_raw(seed, gid, episode, kDomainReset, j);
      set_entity(2 * j, bits_to_pos<T>(r.x), bits_to_pos<T>(r.y));
      set_entity(2 * j + 1, bits_to_pos<T>(r.z), bits_to_pos<T>(r.w));
    }
#pragma unroll
    for (int i = 0; i < kTrN; ++i) vx[i] = vy[i] = (T)0;
    const uint4 t0 = philox_raw(seed, gid, episode, kDomainGoal, 0), t1 = philox_raw(seed, gid, episode, kDomainGoal, 1);
    flags = (t0.x >> 31) | ((t0.y >> 31) << 1) | ((t0.z >> 31) << 2) | ((t0.w >> 31) << 3) | ((t1.x >> 31) << 4) |
            ((t1.y >> 31) << 5) | (0x3Fu << 6);
  }
  __device__ __forceinline__ void set_entity(int e, T x, T y) {
    if (e < kTrN) { px[e] = x; py[e] = y; }
    else { tx[e - kTrN] = x * (T)0.95; ty[e - kTrN] = y * (T)0.95; }
  }

  // _set_action (one-hot branch, sensitivity = accel) + World.step up to integrate_state
  __device__ __forceinline__ void physics(const int *au, const EnvState<T> &s) {
    const bool has_accel = s.accel >= (T)0;
    const T sens = has_accel ? s.accel : (T)5.0;
    T fx[kTrN], fy[kTrN];
#pragma unroll
    for (int i = 0; i < kTrN; ++i) {
      const int a = au[i];
      const T u0 = ((T)0 + ((a == 1 ? (T)1 : (T)0) - (a == 2 ? (T)1 : (T)0))) * sens;
      const T u1 = ((T)0 + ((a == 3 ? (T)1 : (T)0) - (a == 4 ? (T)1 : (T)0))) * sens;
      const T mass = collector(i) ? (T)1.0 : (T)2.25;
      const T k = has_accel ? mass * s.accel : mass;  // apply_action_force: (mass * accel) * action.u
      fx[i] = k * u0;
      fy[i] = k * u1;
    }
#pragma unroll
    for (int a = 0; a < kTrN; ++a) {
#pragma unroll
      for (int b = a + 1; b < kTrN; ++b) {
        const T dist_min = (collector(a) ? (T)0.05 : (T)0.075) + (collector(b) ? (T)0.05 : (T)0.075);
        const T cut = collector(b) ? s.tr_cut[0] : (collector(a) ? s.tr_cut[1] : s.tr_cut[2]);
        const T dx = px[a] - px[b], dy = py[a] - py[b];
        const T d2 = sq2<T>(dx, dy);
        if (std::is_same<T, float>::value || !(d2 >= cut)) {
          T gx, gy;
          contact_force<T>(dx, dy, d2, dist_min, gx, gy);
          if (collector(a) != collector(b)) {  // a collector (mass 1), b deposit (mass 2.25): force_ratio = m_b / m_a
            const T ratio = (T)2.25, inv = (T)(1.0 / 2.25);
            fx[a] = ratio * gx + fx[a]; fy[a] = ratio * gy + fy[a];
            fx[b] = -inv * gx + fx[b]; fy[b] = -inv * gy + fy[b];
          } else {
            fx[a] = gx + fx[a]; fy[a] = gy + fy[a];
            fx[b] = -gx + fx[b]; fy[b] = -gy + fy[b];
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kTrN; ++i) {
      if (collector(i)) {
        integrate_agent<T>(px[i], py[i], vx[i], vy[i], fx[i], fy[i], s.max_speed);
      } else if (std::is_same<T, float>::value) {
        integrate_agent<T>(px[i], py[i], vx[i], vy[i], fx[i] * (T)(1.0 / 2.25), fy[i] * (T)(1.0 / 2.25), s.max_speed);
      } else {
        integrate_agent<T>(px[i], py[i], vx[i], vy[i], fx[i] / (T)2.25, fy[i] / (T)2.25, s.max_speed);
      }
    }
  }

  // Observation of agent i -> row[0..30); key[l] = distance (fp64 build) or squared distance (fp32) to treasure l
  __device__ __forceinline__ void obs_row(int i, T *row, T (&key)[kTrL]) const {
    T dxl[kTrL], dyl[kTrL];
    int rank[kTrL];
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      dxl[l] = tx[l] - px[i];
      dyl[l] = ty[l] - py[i];
      const T d2 = sq2<T>(dxl[l], dyl[l]);
      key[l] = std::is_same<T, float>::value ? d2 : sqrt(d2);  // sorted(zip(cached_dist_mag, index)): ties by index
      rank[l] = 0;
    }
#pragma unroll
    for (int l = 0; l < kTrL; ++l)
#pragma unroll
      for (int m = l + 1; m < kTrL; ++m) {
        const bool gt = key[l] > key[m];
        rank[l] += gt ? 1 : 0;
        rank[m] += gt ? 0 : 1;
      }
    const int h = collector(i) ? tr_hold(flags, i) : -1;
    row[0] = px[i]; row[1] = py[i]; row[2] = vx[i]; row[3] = vy[i];
    row[4] = h == 0 ? (T)1 : (T)0;
    row[5] = h == 1 ? (T)1 : (T)0;
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      T *dst = row + 6 + 4 * rank[l];
      const int ty_ = tr_type(flags, l);
      dst[0] = dxl[l]; dst[1] = dyl[l];
      dst[2] = ty_ == 0 ? (T)1 : (T)0;
      dst[3] = ty_ == 1 ? (T)1 : (T)0;
    }
  }
};

// Observation rows of the warp's envs -> obs[b][8][30]; also returns, for the collectors of this thread's env, the
// collector-treasure contact bits (bit i * 6 + l) and the distance key of the nearest treasure.
template <typename T>
__device__ __forceinline__ void tr_emit_obs(const TreasureEnv<T> &e, const EnvState<T> &s, T *obs, int64_t b0, int lane,
                                            bool full, bool active, T *st_obs, uint64_t &ct, T (&near)[kTrC]) {
  ct = 0ull;
  const bool staged = full && obs != nullptr;
#pragma unroll
  for (int c = 0; c < kTrN / 2; ++c) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = 2 * c + k;
      T key[kTrL];
      // partial warps write their rows straight to global memory; lanes without a row to write use the staging
      // buffer as a sink (the distance keys are needed either way)
      T *row = (staged || obs == nullptr || !active) ? st_obs + lane * kTrStride + k * kTrD : obs + (b0 + lane) * kTrR + i * kTrD;
      e.obs_row(i, row, key);
      if (i < kTrC) {
        T m = key[0];
#pragma unroll
        for (int l = 0; l < kTrL; ++l) {
          m = key[l] < m ? key[l] : m;
          const bool hit = std::is_same<T, float>::value ? key[l] < s.tr_t2[2] : key[l] < (T)(0.05 + 0.025);
          ct |= hit ? (1ull << (i * 6 + l)) : 0ull;
        }
        near[i] = m;
      }
    }
    if (staged) {
      __syncwarp();
      T *dst = obs + b0 * kTrR + c * kTrChunk;
      for (int idx = lane; idx < 32 * kTrChunk; idx += 32) {
        const int env = idx / kTrChunk, j = idx - env * kTrChunk;
        dst[(int64_t)env * kTrR + j] = st_obs[env * kTrStride + j];
      }
      __syncwarp();
    }
  }
}

// MODE 0: env.step   1: env.reset (masked / timed-out envs) + observation   2: observation only
template <typename T, int MODE>
__global__ void __launch_bounds__(kStepThreads)
    k_treasure(EnvState<T> s, const int32_t *__restrict__ act_u, const uint8_t *__restrict__ mask, int auto_len,
               T *__restrict__ obs, T *__restrict__ rew, uint8_t *__restrict__ done, int32_t *__restrict__ info_i) {
  extern __shared__ __align__(128) unsigned char smem[];
  using TL = TrLayout<T>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b0 = (int64_t)blockIdx.x * kStepThreads + warp * 32;
  const int64_t b = b0 + lane;
  const bool active = b < s.B;
  const bool full = b0 + 32 <= s.B;
  T *st_obs = reinterpret_cast<T *>(smem + warp * TL::kWarpBytes);
  T *st_rew = st_obs + TL::kObsElems;
  TreasureEnv<T> e;
  uint32_t ep = 0;
  int tstep = 0;
  if (MODE == 1) {
    double ret = 0.0, n_ep = 0.0, n_steps = 0.0;
    bool doit = false;
    if (active) {
      doit = (mask == nullptr || mask[b] != 0) && (auto_len <= 0 || s.tstep[b] >= auto_len);
      if (doit) {
        ep = s.episode[b] + 1u;
        s.episode[b] = ep;
        e.reset(s.seed, (uint64_t)(s.gid0 + b), ep);
        e.store_agents(s, b);
        e.store_treasures(s, b, 0x3Fu);
        const int t = s.tstep[b];
        if (s.track && t > 0) { ret = (double)s.ep_ret[b]; n_ep = 1.0; n_steps = (double)t; }
        s.tstep[b] = 0;
        s.ep_ret[b] = (T)0;
      } else if (obs != nullptr) {
        e.load(s, b);
      }
    }
    if (s.track) fold_stats(s.stats, ret, n_ep, n_steps);
  } else if (active) {
    e.load(s, b);
  }
  if (MODE == 0 && active) {
    int au[kTrN];
#pragma unroll
    for (int i = 0; i < kTrN; ++i) au[i] = act_u[b * kTrN + i];
    e.physics(au, s);
    e.store_agents(s, b);
    ep = s.episode[b];
    tstep = s.tstep[b];
  }
  uint64_t ct;
  T near[kTrC];
  if (MODE != 0 && obs == nullptr) return;
  tr_emit_obs<T>(e, s, obs, b0, lane, full, active, st_obs, ct, near);
  if (MODE != 0) return;

  // ---- rewards (taken BEFORE post_step) ----
  T r[kTrN];
  int bench[kTrN];
  uint32_t moved = 0u;
  if (active) {
    const uint32_t f = e.flags;
    int hold[kTrC];
#pragma unroll
    for (int i = 0; i < kTrC; ++i) hold[i] = tr_hold(f, i);
    // collector-collector contacts, collector-deposit distances
    int ncc[kTrC];
#pragma unroll
    for (int i = 0; i < kTrC; ++i) ncc[i] = 0;
#pragma unroll
    for (int i = 0; i < kTrC; ++i)
#pragma unroll
      for (int j = i + 1; j < kTrC; ++j) {
        const T d2 = sq2<T>(e.px[i] - e.px[j], e.py[i] - e.py[j]);
        const bool hit = d2 < s.tr_t2[0];
        ncc[i] += hit ? 1 : 0;
        ncc[j] += hit ? 1 : 0;
      }
    T dcd[kTrC][2];   // distance collector i - deposit d
    uint32_t cd = 0u;  // bit i * 2 + d: in contact
#pragma unroll
    for (int i = 0; i < kTrC; ++i)
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const T d2 = sq2<T>(e.px[i] - e.px[kTrC + d], e.py[i] - e.py[kTrC + d]);
        dcd[i][d] = sqrt(d2);
        cd |= (d2 < s.tr_t2[1]) ? (1u << (i * 2 + d)) : 0u;
      }
    int g_dep = 0, g_col = 0;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      int n = 0;
#pragma unroll
      for (int i = 0; i < kTrC; ++i) n += (hold[i] == d && ((cd >> (i * 2 + d)) & 1u)) ? 1 : 0;
      g_dep += 5 * n;
    }
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      int n = 0;
#pragma unroll
      for (int i = 0; i < kTrC; ++i) n += (hold[i] < 0 && ((ct >> (i * 6 + l)) & 1ull)) ? 1 : 0;
      g_col += 5 * n;
    }
    const T glob = (T)(g_dep + g_col);
#pragma unroll
    for (int i = 0; i < kTrC; ++i) {
      const T nearest = std::is_same<T, float>::value ? sqrt(near[i]) : near[i];
      const T shaped = hold[i] < 0 ? nearest : (hold[i] == 0 ? dcd[i][0] : dcd[i][1]);
      T rr = (T)(-5 * ncc[i]);
      rr -= (T)0.1 * shaped;
      r[i] = rr + glob;
      const bool at_dep = hold[i] >= 0 && ((cd >> (i * 2 + (hold[i] > 0 ? 1 : 0))) & 1u);
      const bool at_tr = hold[i] < 0 && ((ct >> (i * 6)) & 0x3Full) != 0ull;
      bench[i] = (at_dep || at_tr) ? 1 : 0;
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      bool any = false;
      T m = (T)0;
#pragma unroll
      for (int i = 0; i < kTrC; ++i)
        if (hold[i] == d) { m = any ? (dcd[i][d] < m ? dcd[i][d] : m) : dcd[i][d]; any = true; }
      if (!any) {  // mean offset of the seven other agents
        T sx = (T)0, sy = (T)0;
#pragma unroll
        for (int j = 0; j < kTrN; ++j)
          if (j != kTrC + d) { sx += e.px[j] - e.px[kTrC + d]; sy += e.py[j] - e.py[kTrC + d]; }
        sx = sx / (T)7; sy = sy / (T)7;
        m = sqrt(sq2<T>(sx, sy));
      }
      T rr = (T)0;
      rr -= (T)0.1 * m;
      r[kTrC + d] = rr + glob;
      bench[kTrC + d] = 0;
    }

    // ---- Scenario.post_step: pick-up, respawn of the treasures collected one step earlier, deposit ----
    uint32_t nf = f;
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      if (tr_alive(f, l)) {
        bool taken = false;
#pragma unroll
        for (int i = 0; i < kTrC; ++i) {
          const bool can = !taken && (((nf >> (12 + 2 * i)) & 3u) == 0u) && ((ct >> (i * 6 + l)) & 1ull);
          if (can) {
            taken = true;
            nf |= (uint32_t)(1 + tr_type(f, l)) << (12 + 2 * i);
          }
        }
        if (taken) {
          nf &= ~(1u << (6 + l));
          e.tx[l] = (T)-999; e.ty[l] = (T)-999;
          moved |= 1u << l;
        }
      } else {
        const uint4 q = philox_raw(s.seed, (uint64_t)(s.gid0 + b), ep, kDomainRespawn, ((uint32_t)(tstep & 0xFFF) << 4) | (uint32_t)l);
        // respawn_prob = 1.0: the probability draw (q.w) always passes
        e.tx[l] = bits_to_pos<T>(q.x) * (T)0.95;
        e.ty[l] = bits_to_pos<T>(q.y) * (T)0.95;
        nf = (nf & ~(1u << l)) | ((q.z >> 31) << l) | (1u << (6 + l));
        moved |= 1u << l;
      }
    }
#pragma unroll
    for (int i = 0; i < kTrC; ++i) {
      const int h = (int)((nf >> (12 + 2 * i)) & 3u) - 1;
      if (h >= 0 && ((cd >> (i * 2 + (h > 0 ? 1 : 0))) & 1u)) nf &= ~(3u << (12 + 2 * i));
    }
    e.flags = nf;
    e.store_treasures(s, b, moved);
    s.tstep[b] = tstep + 1;
    if (s.track) {
      T sum = (T)0;
#pragma unroll
      for (int i = 0; i < kTrN; ++i) sum += r[i];
      s.ep_ret[b] += sum;
    }
    if (info_i != nullptr) {
#pragma unroll
      for (int i = 0; i < kTrN; ++i) info_i[b * (kTrN + 1) + i] = bench[i];
      info_i[b * (kTrN + 1) + kTrN] = 0;
    }
  }
  if (done != nullptr) {
    if (full) {
      uint32_t *d32 = reinterpret_cast<uint32_t *>(done + b0 * kTrN);  // 256 bytes per warp
      for (int k = lane; k < 8 * kTrN; k += 32) d32[k] = 0u;
    } else if (active) {
#pragma unroll
      for (int i = 0; i < kTrN; ++i) done[b * kTrN + i] = 0;
    }
  }
  if (rew != nullptr) {
    if (full) {
#pragma unroll
      for (int i = 0; i < kTrN; ++i) st_rew[lane * kTrN + i] = r[i];
      __syncwarp();
      for (int idx = lane; idx < 32 * kTrN; idx += 32) rew[b0 * kTrN + idx] = st_rew[idx];
    } else if (active) {
#pragma unroll
      for (int i = 0; i < kTrN; ++i) rew[b * kTrN + i] = r[i];
    }
  }
}

}  // namespace mpe
