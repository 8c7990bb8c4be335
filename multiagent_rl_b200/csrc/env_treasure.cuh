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
// Mapping: 4 LANES PER ENV, 8 envs per warp.  Lane q of an env owns agents 2q and 2q + 1 (lanes 0..2: the six
// collectors, lane 3: the two deposits - a lane's two agents share mass, size and role) and, for q < 3, treasures 2q
// and 2q + 1; everything another lane needs travels by warp shuffle:
//   * forces, fp32: most pairs ONCE - the pair inside the lane's block once (applied to both agents), the 2 x 2 pairs
//     against block q + 1 evaluated here and handed (negated) to that block's lane through the same shuffles that bring
//     block q - 1's, the opposite block q + 2 from both sides: 9 contact evaluations per lane for 28 pairs per env.
//     fp64: every agent adds its 7 contacts in the order of the other agent's index (upstream's (a, b) lexicographic
//     pair order seen from that agent); the mass ratio is applied from the lane's own side (f_a = r f and
//     f_b = -(1/r) f negate exactly, so either side computes the bits upstream computes);
//   * observation: 6 treasure positions by shuffle; 15 comparisons give every treasure its rank in an agent's sorted
//     list and its entry is stored at the rank's offset of the agent's row in shared memory; the warp's 64 rows are
//     one contiguous 7,680 B span of obs[b][8][30] and leave with ONE cp.async.bulk issued by lane 0;
//   * rewards: every lane from its own agents' distances; the global term is one REDUX over the env's 4 lanes;
//   * post_step (pick-up / respawn / deposit): replicated integer logic on the env's state word behind warp-uniform
//     early-outs (a pick-up or a deposit happens in a few per cent of the env steps).
// Per-env integer state is one word (EnvState::goal):
//   bit l        type of treasure l (two types = the two deposits)
//   bit 6 + l    treasure l is alive (a collected treasure sits at (-999, -999) until it respawns one step later)
//   bits 12+2i   what collector i holds: 0 = nothing, 1 + type otherwise
// History (ncu digests in profiles/r2_ncu_treasure.txt): thread per env (168 registers, 12 warps per SM, latency
// bound) 0.42 -> 0.64 of HBM; one lane per agent (instruction bound: 929 warp instructions per 4 envs, the per-lane
// setup / decoding / reductions replicated 8 times per env) 0.70.
#pragma once
#include "env_core.cuh"

namespace mpe {

constexpr int kTrN = 8, kTrC = 6, kTrL = 6, kTrD = 30, kTrR = kTrN * kTrD;
constexpr int kTrG = 4, kTrA = kTrN / kTrG;  // lanes per env, agents (and treasures) per lane
constexpr int kTrEpw = 32 / kTrG;            // envs per warp

template <typename T>
struct TrLayout {
  static constexpr int kWarpBytes = (kTrEpw * kTrR * (int)sizeof(T) + 127) / 128 * 128;  // 64 rows of 30 values
  static constexpr int kBlockBytes = kWarpBytes * (kStepThreads / 32);
  static constexpr int kEnvsPerBlock = kTrEpw * (kStepThreads / 32);
};

__device__ __forceinline__ int tr_type(uint32_t f, int l) { return (int)((f >> l) & 1u); }
__device__ __forceinline__ bool tr_alive(uint32_t f, int l) { return ((f >> (6 + l)) & 1u) != 0u; }

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

template <typename T>
__device__ __forceinline__ T tr_shfl(T v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// MODE 0: env.step   1: env.reset (masked / timed-out envs) + observation   2: observation only
template <typename T, int MODE>
__global__ void __launch_bounds__(kStepThreads, std::is_same<T, float>::value ? 6 : 1)
    k_treasure(EnvState<T> s, const int32_t *__restrict__ act_u, const uint8_t *__restrict__ mask, int auto_len,
               T *__restrict__ obs, T *__restrict__ rew, uint8_t *__restrict__ done, int32_t *__restrict__ info_i) {
  extern __shared__ __align__(128) unsigned char smem[];
  using TL = TrLayout<T>;
  constexpr bool kF32 = std::is_same<T, float>::value;
  constexpr int A = kTrA, G = kTrG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int el = lane / G, q = lane - el * G, base = el * G;  // env within the warp, lane within the env, its first lane
  const int64_t b0 = ((int64_t)blockIdx.x * (kStepThreads / 32) + warp) * kTrEpw;
  const bool active = b0 + el < s.B;
  const int64_t b = active ? b0 + el : s.B - 1;  // lanes beyond the batch shadow the last env; their stores are predicated off
  const bool full = b0 + kTrEpw <= s.B;
  const bool coll = q < kTrC / A;                // lanes 0..2: collectors (mass 1, size 0.05); lane 3: deposits (2.25, 0.075)
  const bool has_tr = q < kTrL / A;              // this lane also keeps treasures 2q, 2q + 1
  const int qt = has_tr ? q : 0;                 // (lane 3 shadows lane 0's treasures: loads stay unpredicated, never stored)
  T *st_obs = reinterpret_cast<T *>(smem + warp * TL::kWarpBytes);
  T *pv_ptr[A], *lm_ptr[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    pv_ptr[k] = s.pv + ((int64_t)(q * A + k) * s.B + b) * 4;
    lm_ptr[k] = s.lm + ((int64_t)(qt * A + k) * s.B + b) * 2;
  }

  T px[A], py[A], vx[A], vy[A], tx[A], ty[A];
  uint32_t f;
  uint32_t ep = 0;
  int tstep = 0;
  if (MODE == 1) {
    double ret = 0.0, n_ep = 0.0, n_steps = 0.0;
    const int t_old = s.tstep[b];
    const uint32_t ep_old = s.episode[b];
    const bool doit = (mask == nullptr || mask[b] != 0) && (auto_len <= 0 || t_old >= auto_len);
    __syncwarp();  // every lane of an env has read the counters before lane 0 of the env rewrites them
    if (doit) {
      // Scenario.reset_world: agents ~ U[-1,1)^2, then per treasure a type and a position ~ 0.95 U[-1,1)^2 (entity e
      // draws from Philox block e >> 1: this lane's two agents are one block, its two treasures another); nobody holds
      ep = ep_old + 1u;
      const uint64_t gid = (uint64_t)(s.gid0 + b);
      const uint4 ra = philox_raw(s.seed, gid, ep, kDomainReset, q);
      px[0] = bits_to_pos<T>(ra.x); py[0] = bits_to_pos<T>(ra.y); px[1] = bits_to_pos<T>(ra.z); py[1] = bits_to_pos<T>(ra.w);
      const uint4 rt = philox_raw(s.seed, gid, ep, kDomainReset, kTrN / 2 + qt);
      tx[0] = bits_to_pos<T>(rt.x) * (T)0.95; ty[0] = bits_to_pos<T>(rt.y) * (T)0.95;
      tx[1] = bits_to_pos<T>(rt.z) * (T)0.95; ty[1] = bits_to_pos<T>(rt.w) * (T)0.95;
      const uint4 t0 = philox_raw(s.seed, gid, ep, kDomainGoal, 0), t1 = philox_raw(s.seed, gid, ep, kDomainGoal, 1);
      f = (t0.x >> 31) | ((t0.y >> 31) << 1) | ((t0.z >> 31) << 2) | ((t0.w >> 31) << 3) | ((t1.x >> 31) << 4) |
          ((t1.y >> 31) << 5) | (0x3Fu << 6);
#pragma unroll
      for (int k = 0; k < A; ++k) {
        vx[k] = vy[k] = (T)0;
        if (active) {
          st4(pv_ptr[k], Vec4<T>{px[k], py[k], vx[k], vy[k]});
          if (has_tr) st2(lm_ptr[k], Vec2<T>{tx[k], ty[k]});
        }
      }
      if (active && q == 0) {
        if (s.track && t_old > 0) { ret = (double)s.ep_ret[b]; n_ep = 1.0; n_steps = (double)t_old; }
        s.goal[b] = (int32_t)f;
        s.episode[b] = ep;
        s.tstep[b] = 0;
        s.ep_ret[b] = (T)0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const Vec4<T> v = ld4(pv_ptr[k]);
        const Vec2<T> t = ld2(lm_ptr[k]);
        px[k] = v.x; py[k] = v.y; vx[k] = v.z; vy[k] = v.w; tx[k] = t.x; ty[k] = t.y;
      }
      f = (uint32_t)s.goal[b];
    }
    if (s.track) fold_stats(s.stats, ret, n_ep, n_steps);
    if (obs == nullptr) return;
  } else {
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const Vec4<T> v = ld4(pv_ptr[k]);
      const Vec2<T> t = ld2(lm_ptr[k]);
      px[k] = v.x; py[k] = v.y; vx[k] = v.z; vy[k] = v.w; tx[k] = t.x; ty[k] = t.y;
    }
    f = (uint32_t)s.goal[b];
  }

  const T size = coll ? (T)0.05 : (T)0.075;
  if (MODE == 0) {
    // ---- _set_action (one-hot branch, sensitivity = accel) + World.step up to integrate_state ----
    int a[A];
#pragma unroll
    for (int k = 0; k < A; ++k) a[k] = act_u[b * kTrN + q * A + k];
    ep = s.episode[b];
    tstep = s.tstep[b];
    const bool has_accel = s.accel >= (T)0;
    const T sens = has_accel ? s.accel : (T)5.0;
    const T mass = coll ? (T)1.0 : (T)2.25;
    const T kf = has_accel ? mass * s.accel : mass;  // apply_action_force: (mass * accel) * action.u
    T fx[A], fy[A];
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const T u0 = ((T)0 + ((a[k] == 1 ? (T)1 : (T)0) - (a[k] == 2 ? (T)1 : (T)0))) * sens;
      const T u1 = ((T)0 + ((a[k] == 3 ? (T)1 : (T)0) - (a[k] == 4 ? (T)1 : (T)0))) * sens;
      fx[k] = kf * u0;
      fy[k] = kf * u1;
    }
    // force_ratio seen from this lane's side against a lane of role `oc`: r = m_other / m_own for the lower index,
    // 1 / (m_own / m_other) for the higher one - the same number either way: 2.25 (collector against deposit), 1 / 2.25
    auto ratio = [&](bool oc) { return (coll == oc) ? (T)1 : (coll ? (T)2.25 : (T)(1.0 / 2.25)); };
    if constexpr (kF32) {
      {  // the pair inside the block: once, applied to both agents (same mass: ratio 1)
        const T dx = px[0] - px[1], dy = py[0] - py[1];
        T gx, gy;
        contact_force<T>(dx, dy, sq2<T>(dx, dy), size + size, gx, gy);
        fx[0] += gx; fy[0] += gy;
        fx[1] -= gx; fy[1] -= gy;
      }
#pragma unroll
      for (int d = 1; 2 * d <= G; ++d) {
        const int pb = (q + d) & (G - 1), rb = (q - d) & (G - 1);
        const bool pc = pb < kTrC / A;
        T ox[A], oy[A], gx[A][A], gy[A][A];
#pragma unroll
        for (int m = 0; m < A; ++m) { ox[m] = tr_shfl(px[m], base + pb); oy[m] = tr_shfl(py[m], base + pb); }
        const T dmin = size + (pc ? (T)0.05 : (T)0.075), sc = ratio(pc);
#pragma unroll
        for (int k = 0; k < A; ++k)
#pragma unroll
          for (int m = 0; m < A; ++m) {
            const T dx = px[k] - ox[m], dy = py[k] - oy[m];
            contact_force<T>(dx, dy, sq2<T>(dx, dy), dmin, gx[k][m], gy[k][m]);
            fx[k] = fmaf(sc, gx[k][m], fx[k]);
            fy[k] = fmaf(sc, gy[k][m], fy[k]);
          }
        if (2 * d < G) {  // block q - d evaluated its agents k against MY agents m: mine is the negative, my own ratio
          const T s2 = ratio(rb < kTrC / A);
#pragma unroll
          for (int k = 0; k < A; ++k)
#pragma unroll
            for (int m = 0; m < A; ++m) {
              fx[m] = fmaf(-s2, tr_shfl(gx[k][m], base + rb), fx[m]);
              fy[m] = fmaf(-s2, tr_shfl(gy[k][m], base + rb), fy[m]);
            }
        }
      }
    } else {
      T qx[kTrN], qy[kTrN];  // every agent of the env (all shuffles before any predicated code)
#pragma unroll
      for (int j = 0; j < kTrN; ++j) { qx[j] = tr_shfl(px[j % A], base + j / A); qy[j] = tr_shfl(py[j % A], base + j / A); }
#pragma unroll
      for (int j = 0; j < kTrN; ++j) {
        const bool jc = j < kTrC;
        const T dist_min = size + (jc ? (T)0.05 : (T)0.075), scale = ratio(jc);
        const T cut = (coll && jc) ? s.tr_cut[0] : ((coll || jc) ? s.tr_cut[1] : s.tr_cut[2]);
#pragma unroll
        for (int k = 0; k < A; ++k) {
          const T dx = px[k] - qx[j], dy = py[k] - qy[j];
          const T d2 = sq2<T>(dx, dy);
          if (j != q * A + k && !(d2 >= cut)) {  // upstream's order (other agent's index ascending), far pairs skipped
            T gx, gy;
            contact_force<T>(dx, dy, d2, dist_min, gx, gy);
            fx[k] = scale * gx + fx[k];
            fy[k] = scale * gy + fy[k];
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < A; ++k) {
      if constexpr (kF32) {
        // v = 0.75 v + (f / m) dt; the max_speed clip scales by max_speed * rsqrt(|v|^2) (MUFU) instead of an IEEE division
        const float im = coll ? 0.1f : (float)(0.1 / 2.25);
        float vxi = fmaf(fx[k], im, vx[k] * 0.75f), vyi = fmaf(fy[k], im, vy[k] * 0.75f);
        if (s.max_speed >= 0.0f) {
          const float s2 = fmaf(vxi, vxi, vyi * vyi);
          float rs;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s2));
          const float sc = s2 > s.max_speed * s.max_speed ? s.max_speed * rs : 1.0f;
          vxi *= sc; vyi *= sc;
        }
        vx[k] = vxi; vy[k] = vyi;
        px[k] = fmaf(vxi, 0.1f, px[k]);
        py[k] = fmaf(vyi, 0.1f, py[k]);
      } else {
        integrate_agent<T>(px[k], py[k], vx[k], vy[k], coll ? fx[k] : fx[k] / (T)2.25, coll ? fy[k] : fy[k] / (T)2.25, s.max_speed);
      }
      if (active) st4(pv_ptr[k], Vec4<T>{px[k], py[k], vx[k], vy[k]});
    }
  }

  // ---- observation rows of the lane's agents (experiments/scenarios.py:95-121) ----
  int hold_own[A];
  uint32_t ct[A];  // bit l: agent k touches treasure l (collector radius; only read on collector lanes)
  T near_key[A];   // distance (fp64 build) or squared distance (fp32) to the nearest treasure
  {
    T txa[kTrL], tya[kTrL];
#pragma unroll
    for (int l = 0; l < kTrL; ++l) { txa[l] = tr_shfl(tx[l % A], base + l / A); tya[l] = tr_shfl(ty[l % A], base + l / A); }
#pragma unroll
    for (int k = 0; k < A; ++k) {
      hold_own[k] = coll ? (int)((f >> (12 + 2 * (q * A + k))) & 3u) - 1 : -1;
      T key[kTrL], dxl[kTrL], dyl[kTrL];
      int rank[kTrL];
      ct[k] = 0u;
#pragma unroll
      for (int l = 0; l < kTrL; ++l) {
        dxl[l] = txa[l] - px[k];
        dyl[l] = tya[l] - py[k];
        const T d2 = sq2<T>(dxl[l], dyl[l]);
        key[l] = kF32 ? d2 : sqrt(d2);  // sorted(zip(cached_dist_mag, index)): ties by index
        rank[l] = l;
        const bool hit = kF32 ? key[l] < s.tr_t2[2] : key[l] < (T)(0.05 + 0.025);
        ct[k] |= hit ? (1u << l) : 0u;
      }
      near_key[k] = key[0];
#pragma unroll
      for (int l = 1; l < kTrL; ++l) near_key[k] = key[l] < near_key[k] ? key[l] : near_key[k];
      // rank[m] = m + #{n > m: key[m] > key[n]} - #{l < m: key[l] > key[m]}
#pragma unroll
      for (int l = 0; l < kTrL; ++l)
#pragma unroll
        for (int m = l + 1; m < kTrL; ++m) {
          const int gt = key[l] > key[m] ? 1 : 0;
          rank[l] += gt;
          rank[m] -= gt;
        }
      if (obs != nullptr) {
        T *row = st_obs + (lane * A + k) * kTrD;  // the env's 8 rows are contiguous: lane q writes rows 2q, 2q + 1
        st2(row, Vec2<T>{px[k], py[k]});
        st2(row + 2, Vec2<T>{vx[k], vy[k]});
        st2(row + 4, Vec2<T>{hold_own[k] == 0 ? (T)1 : (T)0, hold_own[k] == 1 ? (T)1 : (T)0});
#pragma unroll
        for (int l = 0; l < kTrL; ++l) {
          T *dst = row + 6 + 4 * rank[l];
          const bool t1 = tr_type(f, l) != 0;
          st2(dst, Vec2<T>{dxl[l], dyl[l]});
          st2(dst + 2, Vec2<T>{t1 ? (T)0 : (T)1, t1 ? (T)1 : (T)0});
        }
      }
    }
  }
  bool issued = false;
  if (obs != nullptr) {
    T *dst = obs + b0 * kTrR;
    if (full && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {  // the warp's 8 envs: 64 x 30 values, contiguous
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store(dst, st_obs, kTrEpw * kTrR * sizeof(T));
        bulk_commit();
        issued = true;
      }
    } else {  // last warp of a ragged batch / unaligned tensor
      __syncwarp();
      const int n_val = (int)((s.B - b0) < kTrEpw ? (s.B - b0) : kTrEpw) * kTrR;
      for (int i = lane; i < n_val; i += 32) dst[i] = st_obs[i];
    }
  }
  if (MODE != 0) {
    if (issued) bulk_wait_read_all();
    return;
  }

  // ---- rewards (taken BEFORE post_step), from the distances of the lane's agents after the step ----
  // what every collector holds comes from the shared state word: holders of deposit type d as a 6-bit mask
  uint32_t holders[2] = {0u, 0u}, free_c = 0u;
#pragma unroll
  for (int i = 0; i < kTrC; ++i) {
    const uint32_t h = (f >> (12 + 2 * i)) & 3u;
    holders[0] |= (h == 1u) ? (1u << i) : 0u;
    holders[1] |= (h == 2u) ? (1u << i) : 0u;
    free_c |= (h == 0u) ? (1u << i) : 0u;
  }
  T nx[kTrN], ny[kTrN];
#pragma unroll
  for (int j = 0; j < kTrN; ++j) { nx[j] = tr_shfl(px[j % A], base + j / A); ny[j] = tr_shfl(py[j % A], base + j / A); }
  T r[A];
  int bench[A], glob_i = 0;
  uint32_t cdbits[A];  // bit d: agent k in contact with deposit d (collector radius pairing)
  T shaped[A];
  int ncc[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    ncc[k] = coll ? -1 : 0;            // contacts with OTHER collectors (the loop counts the agent's own d2 = 0 as one)
    T d2dep0 = (T)0, d2dep1 = (T)0;    // squared distance to deposit 0 / 1
    T m2hold = (T)3.0e38;              // deposit lane: squared distance to the nearest collector holding type k
    T sx = (T)0, sy = (T)0;            // sum of the offsets of the seven other agents (the own offset is an exact 0)
    cdbits[k] = 0u;
#pragma unroll
    for (int j = 0; j < kTrN; ++j) {
      const T ox = nx[j] - px[k], oy = ny[j] - py[k];
      const T d2 = sq2<T>(px[k] - nx[j], py[k] - ny[j]);
      if (j < kTrC) {
        ncc[k] += d2 < s.tr_t2[0] ? 1 : 0;
        const T cand = ((holders[k] >> j) & 1u) ? d2 : (T)3.0e38;  // deposit lane: agent k IS deposit k
        m2hold = cand < m2hold ? cand : m2hold;
      } else {
        if (j == kTrC) d2dep0 = d2; else d2dep1 = d2;
        cdbits[k] |= (d2 < s.tr_t2[1]) ? (1u << (j - kTrC)) : 0u;
      }
      sx += ox; sy += oy;
    }
    const bool at_dep = hold_own[k] >= 0 && ((cdbits[k] >> (hold_own[k] > 0 ? 1 : 0)) & 1u);
    // global reward: 5 per (deposit, matching holder in contact) + 5 per (treasure, free collector in contact)
    glob_i += coll ? ((hold_own[k] < 0 ? 5 * __popc(ct[k]) : 0) + (at_dep ? 5 : 0)) : 0;
    if (coll) {
      // nearest treasure while holding nothing, else the deposit of the held type (sqrt is monotone: min of squares)
      const T kk = hold_own[k] < 0 ? near_key[k] : (hold_own[k] == 0 ? d2dep0 : d2dep1);
      shaped[k] = (kF32 || hold_own[k] >= 0) ? tr_sqrt<T>(kk) : kk;  // fp64: near_key already is a distance
      bench[k] = (at_dep || (hold_own[k] < 0 && ct[k] != 0u)) ? 1 : 0;
    } else {
      sx = sx / (T)7; sy = sy / (T)7;
      shaped[k] = tr_sqrt<T>(holders[k] != 0u ? m2hold : sq2<T>(sx, sy));
      bench[k] = 0;
    }
  }
  const unsigned env_mask = ((1u << G) - 1u) << base;  // the lanes of this env
  glob_i = __reduce_add_sync(env_mask, glob_i);
  const T glob = (T)glob_i;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    T rr = coll ? (T)(-5 * ncc[k]) : (T)0;
    rr -= (T)0.1 * shaped[k];
    r[k] = rr + glob;
  }

  // ---- Scenario.post_step: pick-up, respawn of the treasures collected one step earlier, deposit ----
  // Every lane runs the same integer logic on the env's state word; the rare parts sit behind warp-uniform tests.
  uint32_t nf = f, taken_mask = 0u;
  const bool touching = coll && ((hold_own[0] < 0 && ct[0] != 0u) || (hold_own[1] < 0 && ct[1] != 0u));
  if (__any_sync(0xffffffffu, touching)) {  // some free collector touches a treasure
    uint32_t cb[kTrC];  // contact bits of every collector, by shuffle
#pragma unroll
    for (int i = 0; i < kTrC; ++i) cb[i] = tr_shfl(ct[i % A], base + i / A);
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {  // branch-free: the warp stays converged
      uint32_t col = 0u;
#pragma unroll
      for (int i = 0; i < kTrC; ++i) col |= ((cb[i] >> l) & 1u) << i;
      const uint32_t cand = tr_alive(f, l) ? (col & free_c) : 0u;
      const bool t = cand != 0u;
      const int i = (__ffs((int)cand) - 1) & 7;
      free_c = t ? (free_c & ~(1u << i)) : free_c;
      nf = t ? ((nf | ((uint32_t)(1 + tr_type(f, l)) << (12 + 2 * i))) & ~(1u << (6 + l))) : nf;
      taken_mask |= t ? (1u << l) : 0u;
    }
  }
  const uint32_t dead = ~(f >> 6) & 0x3Fu;  // collected one step earlier: respawn now (respawn_prob = 1.0: the draw always passes)
  bool moved[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    moved[k] = has_tr && ((taken_mask >> (q * A + k)) & 1u);
    if (moved[k]) { tx[k] = (T)-999; ty[k] = (T)-999; }
  }
  if (dead != 0u) {  // rare
#pragma unroll 1
    for (int l = 0; l < kTrL; ++l) {
      if (!((dead >> l) & 1u)) continue;
      const uint4 w = philox_raw(s.seed, (uint64_t)(s.gid0 + b), ep, kDomainRespawn, ((uint32_t)(tstep & 0xFFF) << 4) | (uint32_t)l);
#pragma unroll
      for (int k = 0; k < A; ++k)
        if (has_tr && l == q * A + k) {
          tx[k] = bits_to_pos<T>(w.x) * (T)0.95;
          ty[k] = bits_to_pos<T>(w.y) * (T)0.95;
          moved[k] = true;
        }
      nf = (nf & ~(1u << l)) | ((w.z >> 31) << l) | (1u << (6 + l));
    }
  }
  {  // deposit: a collector that (now) holds type h and touches deposit h lets go; each lane decides for its own agents
    uint32_t clear = 0u;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const int i = q * A + k;
      const int h = coll ? (int)((nf >> (12 + 2 * i)) & 3u) - 1 : -1;
      clear |= (h >= 0 && ((cdbits[k] >> (h > 0 ? 1 : 0)) & 1u)) ? (3u << (12 + 2 * i)) : 0u;
    }
    if (__any_sync(0xffffffffu, clear != 0u)) nf &= ~__reduce_or_sync(env_mask, clear);
  }
  T team = (T)0;
  if (s.track) {  // team return, added in agent order like the other kernels
#pragma unroll
    for (int j = 0; j < kTrN; ++j) team += tr_shfl(r[j % A], base + j / A);
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const int64_t row = b * kTrN + q * A + k;
      if (moved[k]) st2(lm_ptr[k], Vec2<T>{tx[k], ty[k]});
      if (rew != nullptr) rew[row] = r[k];
      if (done != nullptr) done[row] = 0;  // no done_callback (experiments/scenarios.py:186-190): always False
      if (info_i != nullptr) info_i[b * (kTrN + 1) + q * A + k] = bench[k];
    }
    if (q == 0) {
      s.goal[b] = (int32_t)nf;
      s.tstep[b] = tstep + 1;
      if (s.track) s.ep_ret[b] += team;
      if (info_i != nullptr) info_i[b * (kTrN + 1) + kTrN] = 0;
    }
  }
  if (issued) bulk_wait_read_all();
}

}  // namespace mpe
