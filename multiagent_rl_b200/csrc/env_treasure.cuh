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
// Output path: observation rows are staged two agents at a time in a per-warp shared-memory tile (row stride 68
// values: the 16 B vectors of a row start in different bank groups), written with 8 B stores at the rank-dependent
// offsets of the sorted treasure list, and leave as coalesced 16 B-per-lane warp stores; rewards are one 32 B sector
// per lane.  The body has no divergent region: lanes beyond the batch work on the last env and only their stores are
// predicated off.  (History, ncu in profiles/r2_ncu_treasure.txt: v1 spent half of its 9.5 k warp instructions per
// step in a copy loop with an integer division per element; v2's per-lane cp.async.bulk was serialised over the lanes.)
#pragma once
#include "env_core.cuh"

namespace mpe {

constexpr int kTrN = 8, kTrC = 6, kTrL = 6, kTrD = 30, kTrR = kTrN * kTrD;
constexpr int kTrChunk = 2 * kTrD;      // two agents' rows per flush
constexpr int kTrStride = kTrChunk + 8;  // shared-memory row stride (values): 16 B multiple, 2-way conflicts for 8 B stores
#ifndef MPE_TR_MIN_BLOCKS
#define MPE_TR_MIN_BLOCKS 3
#endif
constexpr int kTrMinBlocks = MPE_TR_MIN_BLOCKS;  // resident CTAs per SM the fp32 kernels are compiled for (register cap)

template <typename T>
struct TrLayout {
  static constexpr int kWarpBytes = (32 * kTrStride * (int)sizeof(T) + 127) / 128 * 128;
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
      if constexpr (std::is_same<T, float>::value) {
        // fp32: v = 0.75 v + (f / m) dt; the max_speed clip scales by max_speed * rsqrt(|v|^2) (MUFU, 2 ulp) instead of
        // dividing by an IEEE square root
        const float im = collector(i) ? 0.1f : (float)(0.1 / 2.25);
        float vxi = fmaf(fx[i], im, vx[i] * 0.75f), vyi = fmaf(fy[i], im, vy[i] * 0.75f);
        if (s.max_speed >= 0.0f) {
          const float s2 = fmaf(vxi, vxi, vyi * vyi);
          float rs;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s2));
          const float sc = s2 > s.max_speed * s.max_speed ? s.max_speed * rs : 1.0f;
          vxi *= sc; vyi *= sc;
        }
        vx[i] = vxi; vy[i] = vyi;
        px[i] = fmaf(vxi, 0.1f, px[i]);
        py[i] = fmaf(vyi, 0.1f, py[i]);
      } else if (collector(i)) {
        integrate_agent<T>(px[i], py[i], vx[i], vy[i], fx[i], fy[i], s.max_speed);
      } else {
        integrate_agent<T>(px[i], py[i], vx[i], vy[i], fx[i] / (T)2.25, fy[i] / (T)2.25, s.max_speed);
      }
    }
  }

  // Observation of agent i -> row[0..30) (8 B aligned); key[l] = distance (fp64 build) or squared distance (fp32)
  // to treasure l
  __device__ __forceinline__ void obs_row(int i, T *row, T (&key)[kTrL], const Vec2<T> (&type1h)[kTrL]) const {
    T dxl[kTrL], dyl[kTrL];
    int rank[kTrL];
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      dxl[l] = tx[l] - px[i];
      dyl[l] = ty[l] - py[i];
      const T d2 = sq2<T>(dxl[l], dyl[l]);
      key[l] = std::is_same<T, float>::value ? d2 : sqrt(d2);  // sorted(zip(cached_dist_mag, index)): ties by index
      rank[l] = l;
    }
    // rank[m] = m + #{n > m: key[m] > key[n]} - #{l < m: key[l] > key[m]}
#pragma unroll
    for (int l = 0; l < kTrL; ++l)
#pragma unroll
      for (int m = l + 1; m < kTrL; ++m) {
        const int gt = key[l] > key[m] ? 1 : 0;
        rank[l] += gt;
        rank[m] -= gt;
      }
    const int h = collector(i) ? tr_hold(flags, i) : -1;
    st2(row, Vec2<T>{px[i], py[i]});
    st2(row + 2, Vec2<T>{vx[i], vy[i]});
    st2(row + 4, Vec2<T>{h == 0 ? (T)1 : (T)0, h == 1 ? (T)1 : (T)0});
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      T *dst = row + 6 + 4 * rank[l];
      st2(dst, Vec2<T>{dxl[l], dyl[l]});
      st2(dst + 2, type1h[l]);
    }
  }
};

template <typename T>
__device__ __forceinline__ T tr_sqrt(T x) {
  if constexpr (std::is_same<T, float>::value) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return sqrt(x);
  }
}

// Observation rows of the warp's 32 envs -> obs[b][8][30], two agents at a time through the warp's staging tile; also
// returns, for the collectors of this lane's env, the collector-treasure contact bits (ct[i] bit l) and the distance
// key of the nearest treasure.  Copy-out of a tile (32 envs x 15 16 B vectors in fp32): round `it` moves envs 2 it and
// 2 it + 1, lanes 0..14 the vectors of the first, lanes 15..29 those of the second, so that every address is a
// per-lane base plus a compile-time offset (16 rounds of LDS.128 + STG.128 with immediate offsets, 30 of 32 lanes
// busy; 240 B contiguous per env).  [A per-lane cp.async.bulk of each env's 240 B was measured first: UBLKCP takes
// uniform operands, so the compiler serialises it over the 32 lanes - a third of the kernel's instructions; a copy
// loop over vector index 32 it + lane spends more on its index arithmetic than on the copies.]
// `n_valid`: envs of this warp that exist (32 except in the last warp).
template <typename T>
__device__ __forceinline__ void tr_emit_obs(const TreasureEnv<T> &e, const EnvState<T> &s, T *obs, int64_t b0, int lane,
                                            int n_valid, T *st_obs, uint32_t (&ct)[kTrC], T (&near)[kTrC]) {
  constexpr int kVec = 16 / (int)sizeof(T);        // values per 16 B vector
  constexpr int kPerEnv = kTrChunk / kVec;          // vectors per env and chunk (15 in fp32, 30 in fp64)
  constexpr int kEnvsPerRound = 32 / kPerEnv;       // 2 (fp32), 1 (fp64)
  const bool vec_ok = (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
  Vec2<T> type1h[kTrL];
#pragma unroll
  for (int l = 0; l < kTrL; ++l) {
    const bool t1 = tr_type(e.flags, l) != 0;
    type1h[l] = Vec2<T>{t1 ? (T)0 : (T)1, t1 ? (T)1 : (T)0};
  }
  T *row = st_obs + lane * kTrStride;
  // copy-out addressing of this lane: env (within a round) and vector slot
  const int sub = lane / kPerEnv, slot = lane - sub * kPerEnv;
  const bool mover = sub < kEnvsPerRound;
  const T *src0 = st_obs + sub * kTrStride + slot * kVec;
  T *dst0 = obs + (b0 + sub) * kTrR + slot * kVec;
#pragma unroll
  for (int c = 0; c < kTrN / 2; ++c) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = 2 * c + k;
      T key[kTrL];
      e.obs_row(i, row + k * kTrD, key, type1h);
      if (i < kTrC) {
        T m = key[0];
        uint32_t bits = 0u;
#pragma unroll
        for (int l = 0; l < kTrL; ++l) {
          m = key[l] < m ? key[l] : m;
          const bool hit = std::is_same<T, float>::value ? key[l] < s.tr_t2[2] : key[l] < (T)(0.05 + 0.025);
          bits |= hit ? (1u << l) : 0u;
        }
        near[i] = m;
        ct[i] = bits;
      }
    }
    __syncwarp();
    if (obs != nullptr) {
      if (vec_ok && n_valid == 32) {
        if (mover) {
#pragma unroll
          for (int it = 0; it < 32 / kEnvsPerRound; ++it) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src0 + it * kEnvsPerRound * kTrStride);
            *reinterpret_cast<uint4 *>(dst0 + (int64_t)it * kEnvsPerRound * kTrR + c * kTrChunk) = v;
          }
        }
      } else if (lane < n_valid) {  // last warp of a ragged batch / unaligned tensor: scalar copy of the lane's own row
        T *dst = obs + (b0 + lane) * kTrR + c * kTrChunk;
        for (int jj = 0; jj < kTrChunk; ++jj) dst[jj] = row[jj];
      }
    }
    __syncwarp();
  }
}

// MODE 0: env.step   1: env.reset (masked / timed-out envs) + observation   2: observation only
template <typename T, int MODE>
__global__ void __launch_bounds__(kStepThreads, std::is_same<T, float>::value ? kTrMinBlocks : 1)
    k_treasure(EnvState<T> s, const int32_t *__restrict__ act_u, const uint8_t *__restrict__ mask, int auto_len,
               T *__restrict__ obs, T *__restrict__ rew, uint8_t *__restrict__ done, int32_t *__restrict__ info_i) {
  extern __shared__ __align__(128) unsigned char smem[];
  using TL = TrLayout<T>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b_raw = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
  const bool active = b_raw < s.B;
  const int64_t b = active ? b_raw : s.B - 1;  // lanes beyond the batch shadow the last env; their stores are predicated off
  T *st_obs = reinterpret_cast<T *>(smem + warp * TL::kWarpBytes);
  TreasureEnv<T> e;
  uint32_t ep = 0;
  int tstep = 0;
  if (MODE == 1) {
    double ret = 0.0, n_ep = 0.0, n_steps = 0.0;
    const int t_old = s.tstep[b];
    const bool doit = (mask == nullptr || mask[b] != 0) && (auto_len <= 0 || t_old >= auto_len);
    if (doit) {
      ep = s.episode[b] + 1u;
      e.reset(s.seed, (uint64_t)(s.gid0 + b), ep);
      if (active) {
        s.episode[b] = ep;
        e.store_agents(s, b);
        e.store_treasures(s, b, 0x3Fu);
        if (s.track && t_old > 0) { ret = (double)s.ep_ret[b]; n_ep = 1.0; n_steps = (double)t_old; }
        s.tstep[b] = 0;
        s.ep_ret[b] = (T)0;
      }
    } else {
      e.load(s, b);
    }
    if (s.track) fold_stats(s.stats, ret, n_ep, n_steps);
    if (obs == nullptr) return;
  } else {
    e.load(s, b);
  }
  if (MODE == 0) {
    int au[kTrN];
    if ((reinterpret_cast<uintptr_t>(act_u) & 15) == 0) {  // the usual case: two 16 B loads per env
      const int4 a0 = *reinterpret_cast<const int4 *>(act_u + b * kTrN), a1 = *reinterpret_cast<const int4 *>(act_u + b * kTrN + 4);
      au[0] = a0.x; au[1] = a0.y; au[2] = a0.z; au[3] = a0.w; au[4] = a1.x; au[5] = a1.y; au[6] = a1.z; au[7] = a1.w;
    } else {
#pragma unroll
      for (int i = 0; i < kTrN; ++i) au[i] = act_u[b * kTrN + i];
    }
    ep = s.episode[b];  // requested before the stores below (no aliasing information: they would be ordered after them)
    tstep = s.tstep[b];
    e.physics(au, s);
    if (active) e.store_agents(s, b);
  }
  uint32_t ct[kTrC];  // collector i touches treasure l: bit l
  T near[kTrC];
  const int64_t b0 = b_raw - lane;
  const int n_valid = (int)((s.B - b0) < 32 ? (s.B - b0) : 32);
  tr_emit_obs<T>(e, s, obs, b0, lane, n_valid, st_obs, ct, near);
  if (MODE != 0) return;

  // ---- rewards (taken BEFORE post_step) ----
  T r[kTrN];
  int bench[kTrN];
  const uint32_t f = e.flags;
  int hold[kTrC];
#pragma unroll
  for (int i = 0; i < kTrC; ++i) hold[i] = tr_hold(f, i);
  int ncc[kTrC];
#pragma unroll
  for (int i = 0; i < kTrC; ++i) ncc[i] = 0;
#pragma unroll
  for (int i = 0; i < kTrC; ++i)
#pragma unroll
    for (int j = i + 1; j < kTrC; ++j) {
      const T d2 = sq2<T>(e.px[i] - e.px[j], e.py[i] - e.py[j]);
      const int hit = d2 < s.tr_t2[0] ? 1 : 0;
      ncc[i] += hit;
      ncc[j] += hit;
    }
  T d2cd[kTrC][2];   // squared distance collector i - deposit d
  uint32_t cd = 0u;  // bit i * 2 + d: in contact
#pragma unroll
  for (int i = 0; i < kTrC; ++i)
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      d2cd[i][d] = sq2<T>(e.px[i] - e.px[kTrC + d], e.py[i] - e.py[kTrC + d]);
      cd |= (d2cd[i][d] < s.tr_t2[1]) ? (1u << (i * 2 + d)) : 0u;
    }
  int g_dep = 0, g_col = 0;
#pragma unroll
  for (int i = 0; i < kTrC; ++i) {
    const bool none = hold[i] < 0;
    g_dep += (!none && ((cd >> (i * 2 + (hold[i] > 0 ? 1 : 0))) & 1u)) ? 5 : 0;
    g_col += none ? 5 * __popc(ct[i]) : 0;
  }
  const T glob = (T)(g_dep + g_col);
#pragma unroll
  for (int i = 0; i < kTrC; ++i) {
    // nearest treasure while holding nothing, else the deposit of the held type (sqrt is monotone: min of squares)
    const T key = hold[i] < 0 ? near[i] : (hold[i] == 0 ? d2cd[i][0] : d2cd[i][1]);
    const T shaped = (std::is_same<T, float>::value || hold[i] >= 0) ? tr_sqrt<T>(key) : key;  // fp64: near[] is a distance
    T rr = (T)(-5 * ncc[i]);
    rr -= (T)0.1 * shaped;
    r[i] = rr + glob;
    const bool at_dep = hold[i] >= 0 && ((cd >> (i * 2 + (hold[i] > 0 ? 1 : 0))) & 1u);
    const bool at_tr = hold[i] < 0 && ct[i] != 0u;
    bench[i] = (at_dep || at_tr) ? 1 : 0;
  }
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    bool any = false;
    T m2 = (T)0;
#pragma unroll
    for (int i = 0; i < kTrC; ++i) {
      const bool h = hold[i] == d;
      m2 = h ? (any ? (d2cd[i][d] < m2 ? d2cd[i][d] : m2) : d2cd[i][d]) : m2;
      any = any || h;
    }
    // no holder of this type: mean offset of the seven other agents
    T sx = (T)0, sy = (T)0;
#pragma unroll
    for (int j = 0; j < kTrN; ++j)
      if (j != kTrC + d) { sx += e.px[j] - e.px[kTrC + d]; sy += e.py[j] - e.py[kTrC + d]; }
    sx = sx / (T)7; sy = sy / (T)7;
    const T m = tr_sqrt<T>(any ? m2 : sq2<T>(sx, sy));
    T rr = (T)0;
    rr -= (T)0.1 * m;
    r[kTrC + d] = rr + glob;
    bench[kTrC + d] = 0;
  }

  // ---- Scenario.post_step: pick-up, respawn of the treasures collected one step earlier, deposit ----
  uint32_t nf = f, moved = 0u;
#pragma unroll
  for (int l = 0; l < kTrL; ++l) {
    bool taken = false;
#pragma unroll
    for (int i = 0; i < kTrC; ++i) {
      const bool can = !taken && (((nf >> (12 + 2 * i)) & 3u) == 0u) && ((ct[i] >> l) & 1u) && tr_alive(f, l);
      nf |= can ? ((uint32_t)(1 + tr_type(f, l)) << (12 + 2 * i)) : 0u;
      taken = taken || can;
    }
    if (taken) {
      nf &= ~(1u << (6 + l));
      e.tx[l] = (T)-999; e.ty[l] = (T)-999;
      moved |= 1u << l;
    }
  }
  uint32_t dead = ~(f >> 6) & 0x3Fu;  // collected one step earlier: respawn now (respawn_prob = 1.0: the draw always passes)
  if (dead != 0u) {                    // rare
#pragma unroll 1
    for (int l = 0; l < kTrL; ++l) {
      if (!((dead >> l) & 1u)) continue;
      const uint4 q = philox_raw(s.seed, (uint64_t)(s.gid0 + b), ep, kDomainRespawn, ((uint32_t)(tstep & 0xFFF) << 4) | (uint32_t)l);
      const T x = bits_to_pos<T>(q.x) * (T)0.95, y = bits_to_pos<T>(q.y) * (T)0.95;
#pragma unroll
      for (int k = 0; k < kTrL; ++k)
        if (k == l) { e.tx[k] = x; e.ty[k] = y; }
      nf = (nf & ~(1u << l)) | ((q.z >> 31) << l) | (1u << (6 + l));
    }
    moved |= dead;
  }
#pragma unroll
  for (int i = 0; i < kTrC; ++i) {
    const int h = (int)((nf >> (12 + 2 * i)) & 3u) - 1;
    if (h >= 0 && ((cd >> (i * 2 + (h > 0 ? 1 : 0))) & 1u)) nf &= ~(3u << (12 + 2 * i));
  }
  e.flags = nf;
  if (active) {
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
    if (done != nullptr) {  // no done_callback (experiments/scenarios.py:186-190): always False
      if ((reinterpret_cast<uintptr_t>(done) & 7) == 0) {
        *reinterpret_cast<uint2 *>(done + b * kTrN) = make_uint2(0u, 0u);
      } else {
#pragma unroll
        for (int i = 0; i < kTrN; ++i) done[b * kTrN + i] = 0;
      }
    }
    if (rew != nullptr) {  // 8 rewards = one 32 B sector (fp32) per lane
      if ((reinterpret_cast<uintptr_t>(rew) & 15) == 0) {
        st4(rew + b * kTrN, Vec4<T>{r[0], r[1], r[2], r[3]});
        st4(rew + b * kTrN + 4, Vec4<T>{r[4], r[5], r[6], r[7]});
      } else {
#pragma unroll
        for (int i = 0; i < kTrN; ++i) rew[b * kTrN + i] = r[i];
      }
    }
  }
}

}  // namespace mpe
