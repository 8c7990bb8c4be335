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
// Mapping: ONE LANE PER AGENT, 8 lanes per env, 4 envs per warp.  Lane q of an env owns agent q (position, velocity)
// and, for q < 6, treasure q; everything another lane needs travels by warp shuffle:
//   * forces: a lane evaluates the 7 contacts of its own agent, adding them in the order of the other agent's index -
//     upstream's (a, b) lexicographic pair order seen from that agent; the mass ratio is applied from the lane's own
//     side (f_a = r f and f_b = -(1/r) f negate exactly, so either side computes the bits upstream computes);
//   * observation: 6 treasure positions by shuffle, 15 comparisons give every treasure its rank in the sorted list and
//     its entry is stored at the rank's offset of the lane's row in shared memory; the warp's 32 rows are one
//     contiguous 3,840 B span of obs[b][8][30] and leave with ONE cp.async.bulk issued by lane 0;
//   * rewards: every lane from its own distances (collector: contacts with collectors, nearest treasure or its deposit;
//     deposit: nearest matching holder or the mean offset of the others); the global term is a 3-step xor reduction;
//   * post_step (pick-up / respawn / deposit): the 6 x 8 contact bits are gathered by 6 shuffles and every lane runs the
//     same integer logic on the env's state word; lane l keeps treasure l.
// Per-env integer state is one word (EnvState::goal):
//   bit l        type of treasure l (two types = the two deposits)
//   bit 6 + l    treasure l is alive (a collected treasure sits at (-999, -999) until it respawns one step later)
//   bits 12+2i   what collector i holds: 0 = nothing, 1 + type otherwise
// History (ncu digests in profiles/r2_ncu_treasure.txt): the thread-per-env versions (every entity in one thread's
// registers, 168 registers, 12 warps per SM) were latency bound at 0.42 -> 0.64 of HBM.
#pragma once
#include "env_core.cuh"

namespace mpe {

constexpr int kTrN = 8, kTrC = 6, kTrL = 6, kTrD = 30, kTrR = kTrN * kTrD;
constexpr int kTrEpw = 32 / kTrN;  // envs per warp

template <typename T>
struct TrLayout {
  static constexpr int kWarpBytes = (32 * kTrD * (int)sizeof(T) + 127) / 128 * 128;  // 32 rows of 30 values
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
__global__ void __launch_bounds__(kStepThreads, std::is_same<T, float>::value ? 10 : 1)  // fp32: 48 registers, 40 warps per SM (8: 62 registers, 77.0 us; 10: 76.2; 12: spills, 77.9)
    k_treasure(EnvState<T> s, const int32_t *__restrict__ act_u, const uint8_t *__restrict__ mask, int auto_len,
               T *__restrict__ obs, T *__restrict__ rew, uint8_t *__restrict__ done, int32_t *__restrict__ info_i) {
  extern __shared__ __align__(128) unsigned char smem[];
  using TL = TrLayout<T>;
  constexpr bool kF32 = std::is_same<T, float>::value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int el = lane >> 3, q = lane & 7, base = el * 8;  // env within the warp, agent, first lane of the env
  const int64_t b0 = ((int64_t)blockIdx.x * (kStepThreads / 32) + warp) * kTrEpw;
  const bool active = b0 + el < s.B;
  const int64_t b = active ? b0 + el : s.B - 1;  // lanes beyond the batch shadow the last env; their stores are predicated off
  const bool full = b0 + kTrEpw <= s.B;
  const bool coll = q < kTrC;                    // collector (mass 1, size 0.05) or deposit (mass 2.25, size 0.075)
  const bool has_tr = q < kTrL;                  // this lane also keeps treasure q
  T *st_obs = reinterpret_cast<T *>(smem + warp * TL::kWarpBytes);
  T *pv_ptr = s.pv + ((int64_t)q * s.B + b) * 4;
  T *lm_ptr = s.lm + ((int64_t)(has_tr ? q : 0) * s.B + b) * 2;

  T px, py, vx, vy, tx = (T)0, ty = (T)0;
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
      // draws from Philox block e >> 1, components by the parity of e); nobody holds anything
      ep = ep_old + 1u;
      const uint64_t gid = (uint64_t)(s.gid0 + b);
      const uint4 ra = philox_raw(s.seed, gid, ep, kDomainReset, q >> 1);
      px = bits_to_pos<T>((q & 1) ? ra.z : ra.x); py = bits_to_pos<T>((q & 1) ? ra.w : ra.y);
      vx = vy = (T)0;
      const uint4 rt = philox_raw(s.seed, gid, ep, kDomainReset, (kTrN + (has_tr ? q : 0)) >> 1);
      tx = bits_to_pos<T>((q & 1) ? rt.z : rt.x) * (T)0.95; ty = bits_to_pos<T>((q & 1) ? rt.w : rt.y) * (T)0.95;
      const uint4 t0 = philox_raw(s.seed, gid, ep, kDomainGoal, 0), t1 = philox_raw(s.seed, gid, ep, kDomainGoal, 1);
      f = (t0.x >> 31) | ((t0.y >> 31) << 1) | ((t0.z >> 31) << 2) | ((t0.w >> 31) << 3) | ((t1.x >> 31) << 4) |
          ((t1.y >> 31) << 5) | (0x3Fu << 6);
      if (active) {
        st4(pv_ptr, Vec4<T>{px, py, vx, vy});
        if (has_tr) st2(lm_ptr, Vec2<T>{tx, ty});
        if (q == 0) {
          if (s.track && t_old > 0) { ret = (double)s.ep_ret[b]; n_ep = 1.0; n_steps = (double)t_old; }
          s.goal[b] = (int32_t)f;
          s.episode[b] = ep;
          s.tstep[b] = 0;
          s.ep_ret[b] = (T)0;
        }
      }
    } else {
      const Vec4<T> v = ld4(pv_ptr);
      px = v.x; py = v.y; vx = v.z; vy = v.w;
      if (has_tr) { const Vec2<T> t = ld2(lm_ptr); tx = t.x; ty = t.y; }
      f = (uint32_t)s.goal[b];
    }
    if (s.track) fold_stats(s.stats, ret, n_ep, n_steps);
    if (obs == nullptr) return;
  } else {
    const Vec4<T> v = ld4(pv_ptr);
    const Vec2<T> t = ld2(lm_ptr);  // lanes 6, 7 read treasure 0 (unused): an unpredicated load is issued up front
    px = v.x; py = v.y; vx = v.z; vy = v.w;
    tx = t.x; ty = t.y;
    f = (uint32_t)s.goal[b];
  }

  if (MODE == 0) {
    // ---- _set_action (one-hot branch, sensitivity = accel) + World.step up to integrate_state ----
    const int a = act_u[b * kTrN + q];
    ep = s.episode[b];
    tstep = s.tstep[b];
    const bool has_accel = s.accel >= (T)0;
    const T sens = has_accel ? s.accel : (T)5.0;
    const T mass = coll ? (T)1.0 : (T)2.25, size = coll ? (T)0.05 : (T)0.075;
    const T u0 = ((T)0 + ((a == 1 ? (T)1 : (T)0) - (a == 2 ? (T)1 : (T)0))) * sens;
    const T u1 = ((T)0 + ((a == 3 ? (T)1 : (T)0) - (a == 4 ? (T)1 : (T)0))) * sens;
    const T k = has_accel ? mass * s.accel : mass;  // apply_action_force: (mass * accel) * action.u
    T fx = k * u0, fy = k * u1;
    if constexpr (kF32) {
      // fp32: every pair ONCE.  The 28 pairs of 8 agents are the 7 xor-rounds q <-> q ^ r; rounds are taken two at a
      // time, (1,3), (2,6), (4,5): with m = lowest bit of the first round, a lane whose bit m is clear evaluates its
      // pair of the first round, a lane whose bit m is set its pair of the second (the second round flips bit m, so
      // exactly one end of every pair qualifies), and each lane receives the force of the one pair it did not
      // evaluate from the lane that did.  Round 7 is evaluated from both ends.  4 contact evaluations per lane instead
      // of 8; the accumulation order differs from upstream's (fp32 tolerance), the fp64 build below keeps it.
      constexpr int kR1[4] = {1, 2, 4, 7}, kR2[4] = {3, 6, 5, 7}, kM[4] = {1, 2, 4, 0};
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        const bool low = (q & kM[sr]) == 0;
        const int p = q ^ (low ? kR1[sr] : kR2[sr]);
        const T ppx = tr_shfl(px, base + p), ppy = tr_shfl(py, base + p);
        const bool pc = p < kTrC;
        const T dx = px - ppx, dy = py - ppy;
        T gx, gy;
        contact_force<T>(dx, dy, sq2<T>(dx, dy), size + (pc ? (T)0.05 : (T)0.075), gx, gy);
        const T sc = (coll == pc) ? (T)1 : (coll ? (T)2.25 : (T)(1.0 / 2.25));
        fx = fmaf(sc, gx, fx);
        fy = fmaf(sc, gy, fy);
        if (sr < 3) {
          const int src = q ^ (low ? kR2[sr] : kR1[sr]);  // evaluated the pair {src, q} with delta = p_src - p_q
          const T hx = tr_shfl(gx, base + src), hy = tr_shfl(gy, base + src);
          const bool sc_c = src < kTrC;
          const T s2 = (coll == sc_c) ? (T)1 : (coll ? (T)2.25 : (T)(1.0 / 2.25));
          fx = fmaf(-s2, hx, fx);
          fy = fmaf(-s2, hy, fy);
        }
      }
    } else {
      T qx[kTrN], qy[kTrN];  // every agent of the env (all shuffles before any predicated code)
#pragma unroll
      for (int j = 0; j < kTrN; ++j) { qx[j] = tr_shfl(px, base + j); qy[j] = tr_shfl(py, base + j); }
#pragma unroll
      for (int j = 0; j < kTrN; ++j) {
        const T pjx = qx[j], pjy = qy[j];
        const bool jc = j < kTrC;
        const T dist_min = size + (jc ? (T)0.05 : (T)0.075);
        const T dx = px - pjx, dy = py - pjy;
        const T d2 = sq2<T>(dx, dy);
        // force_ratio seen from this lane's side: r = m_other / m_own for the lower index, 1 / (m_own / m_other) for
        // the higher one - the same number either way: 2.25 (collector against deposit), 1 / 2.25 (the reverse)
        const T scale = (coll == jc) ? (T)1 : (coll ? (T)2.25 : (T)(1.0 / 2.25));
        const T cut = (coll && jc) ? s.tr_cut[0] : ((coll || jc) ? s.tr_cut[1] : s.tr_cut[2]);
        if (j != q && !(d2 >= cut)) {  // fp64: upstream's order (other agent's index ascending), far pairs skipped
          T gx, gy;
          contact_force<T>(dx, dy, d2, dist_min, gx, gy);
          fx = scale * gx + fx;
          fy = scale * gy + fy;
        }
      }
    }
    if constexpr (kF32) {
      // v = 0.75 v + (f / m) dt; the max_speed clip scales by max_speed * rsqrt(|v|^2) (MUFU) instead of an IEEE division
      const float im = coll ? 0.1f : (float)(0.1 / 2.25);
      float vxi = fmaf(fx, im, vx * 0.75f), vyi = fmaf(fy, im, vy * 0.75f);
      if (s.max_speed >= 0.0f) {
        const float s2 = fmaf(vxi, vxi, vyi * vyi);
        float rs;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s2));
        const float sc = s2 > s.max_speed * s.max_speed ? s.max_speed * rs : 1.0f;
        vxi *= sc; vyi *= sc;
      }
      vx = vxi; vy = vyi;
      px = fmaf(vxi, 0.1f, px);
      py = fmaf(vyi, 0.1f, py);
    } else {
      integrate_agent<T>(px, py, vx, vy, coll ? fx : fx / (T)2.25, coll ? fy : fy / (T)2.25, s.max_speed);
    }
    if (active) st4(pv_ptr, Vec4<T>{px, py, vx, vy});
  }

  // ---- observation row of agent q (experiments/scenarios.py:95-121) ----
  const int hold_own = coll ? (int)((f >> (12 + 2 * q)) & 3u) - 1 : -1;
  uint32_t ct = 0u;  // bit l: this agent touches treasure l (collector radius; only read on collector lanes)
  T near_key;        // distance (fp64 build) or squared distance (fp32) to the nearest treasure
  {
    T key[kTrL], dxl[kTrL], dyl[kTrL];
    int rank[kTrL];
#pragma unroll
    for (int l = 0; l < kTrL; ++l) { dxl[l] = tr_shfl(tx, base + l); dyl[l] = tr_shfl(ty, base + l); }
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {
      dxl[l] = dxl[l] - px;
      dyl[l] = dyl[l] - py;
      const T d2 = sq2<T>(dxl[l], dyl[l]);
      key[l] = kF32 ? d2 : sqrt(d2);  // sorted(zip(cached_dist_mag, index)): ties by index
      rank[l] = l;
      const bool hit = kF32 ? key[l] < s.tr_t2[2] : key[l] < (T)(0.05 + 0.025);
      ct |= hit ? (1u << l) : 0u;
    }
    near_key = key[0];
#pragma unroll
    for (int l = 1; l < kTrL; ++l) near_key = key[l] < near_key ? key[l] : near_key;
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
      T *row = st_obs + lane * kTrD;  // 120 B rows: the 8 B stores of a half-warp start in 16 different bank pairs
      st2(row, Vec2<T>{px, py});
      st2(row + 2, Vec2<T>{vx, vy});
      st2(row + 4, Vec2<T>{hold_own == 0 ? (T)1 : (T)0, hold_own == 1 ? (T)1 : (T)0});
#pragma unroll
      for (int l = 0; l < kTrL; ++l) {
        T *dst = row + 6 + 4 * rank[l];
        const bool t1 = tr_type(f, l) != 0;
        st2(dst, Vec2<T>{dxl[l], dyl[l]});
        st2(dst + 2, Vec2<T>{t1 ? (T)0 : (T)1, t1 ? (T)1 : (T)0});
      }
    }
  }
  bool issued = false;
  if (obs != nullptr) {
    T *dst = obs + b0 * kTrR;
    if (full && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {  // the warp's 4 envs: 32 x 30 values, contiguous
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store(dst, st_obs, 32 * kTrD * sizeof(T));
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

  // ---- rewards (taken BEFORE post_step), from the distances of this lane's agent after the step ----
  // holders of this deposit lane's type among the collectors (bit j), from two ballots of the collectors' own holdings
  const uint32_t holds0 = (__ballot_sync(0xffffffffu, hold_own == 0) >> base) & 0x3Fu;
  const uint32_t holds1 = (__ballot_sync(0xffffffffu, hold_own == 1) >> base) & 0x3Fu;
  const uint32_t holders = q == kTrC ? holds0 : holds1;
  int ncc = coll ? -1 : 0;            // contacts with OTHER collectors (the loop counts the lane's own d2 = 0 as one)
  T d2dep0 = (T)0, d2dep1 = (T)0;     // squared distance to deposit 0 / 1
  T m2hold = (T)3.0e38;               // deposit lanes: squared distance to the nearest collector that holds their type
  T sx = (T)0, sy = (T)0;             // sum of the offsets of the seven other agents (the own offset is an exact 0)
  uint32_t cdbits = 0u;               // bit d: in contact with deposit d (collector radius pairing)
  T nx[kTrN], ny[kTrN];
#pragma unroll
  for (int j = 0; j < kTrN; ++j) { nx[j] = tr_shfl(px, base + j); ny[j] = tr_shfl(py, base + j); }
#pragma unroll
  for (int j = 0; j < kTrN; ++j) {
    const T ox = nx[j] - px, oy = ny[j] - py;
    const T d2 = sq2<T>(px - nx[j], py - ny[j]);
    if (j < kTrC) {
      ncc += d2 < s.tr_t2[0] ? 1 : 0;
      const T cand = ((holders >> j) & 1u) ? d2 : (T)3.0e38;
      m2hold = cand < m2hold ? cand : m2hold;
    } else {
      if (j == kTrC) d2dep0 = d2; else d2dep1 = d2;
      cdbits |= (d2 < s.tr_t2[1]) ? (1u << (j - kTrC)) : 0u;
    }
    sx += ox; sy += oy;
  }
  const bool any_hold = holders != 0u;
  const bool at_dep = hold_own >= 0 && ((cdbits >> (hold_own > 0 ? 1 : 0)) & 1u);
  // global reward: 5 per (deposit, matching holder in contact) + 5 per (treasure, free collector in contact)
  int glob_i = coll ? ((hold_own < 0 ? 5 * __popc(ct) : 0) + (at_dep ? 5 : 0)) : 0;
  const unsigned env_mask = 0xFFu << base;  // the 8 lanes of this env
  glob_i = __reduce_add_sync(env_mask, glob_i);
  const T glob = (T)glob_i;
  T r;
  int bench;
  if (coll) {
    // nearest treasure while holding nothing, else the deposit of the held type (sqrt is monotone: min of squares)
    const T kk = hold_own < 0 ? near_key : (hold_own == 0 ? d2dep0 : d2dep1);
    const T shaped = (kF32 || hold_own >= 0) ? tr_sqrt<T>(kk) : kk;  // fp64: near_key already is a distance
    T rr = (T)(-5 * ncc);
    rr -= (T)0.1 * shaped;
    r = rr + glob;
    bench = (at_dep || (hold_own < 0 && ct != 0u)) ? 1 : 0;
  } else {
    sx = sx / (T)7; sy = sy / (T)7;
    const T m = tr_sqrt<T>(any_hold ? m2hold : sq2<T>(sx, sy));
    T rr = (T)0;
    rr -= (T)0.1 * m;
    r = rr + glob;
    bench = 0;
  }

  // ---- Scenario.post_step: pick-up, respawn of the treasures collected one step earlier, deposit ----
  // Every lane runs the same integer logic on the env's state word.  col[l] = collectors in contact with treasure l
  // (a ballot per treasure, this env's byte of it); pick-up: the lowest-index free collector of col[l].
  uint32_t nf = f, taken_mask = 0u;
  if (__any_sync(0xffffffffu, coll && hold_own < 0 && ct != 0u)) {  // some free collector touches a treasure: rare
    uint32_t free_c = (__ballot_sync(0xffffffffu, coll && hold_own < 0) >> base) & 0x3Fu;
    uint32_t col[kTrL];
#pragma unroll
    for (int l = 0; l < kTrL; ++l) col[l] = (__ballot_sync(0xffffffffu, (ct >> l) & 1u) >> base) & 0x3Fu;
#pragma unroll
    for (int l = 0; l < kTrL; ++l) {  // branch-free: the warp stays converged
      const uint32_t cand = tr_alive(f, l) ? (col[l] & free_c) : 0u;
      const bool t = cand != 0u;
      const int i = (__ffs((int)cand) - 1) & 7;
      free_c = t ? (free_c & ~(1u << i)) : free_c;
      nf = t ? ((nf | ((uint32_t)(1 + tr_type(f, l)) << (12 + 2 * i))) & ~(1u << (6 + l))) : nf;
      taken_mask |= t ? (1u << l) : 0u;
    }
  }
  const uint32_t dead = ~(f >> 6) & 0x3Fu;  // collected one step earlier: respawn now (respawn_prob = 1.0: the draw always passes)
  bool moved = has_tr && ((taken_mask >> q) & 1u);
  if (moved) { tx = (T)-999; ty = (T)-999; }
  if (dead != 0u) {  // rare
#pragma unroll 1
    for (int l = 0; l < kTrL; ++l) {
      if (!((dead >> l) & 1u)) continue;
      const uint4 w = philox_raw(s.seed, (uint64_t)(s.gid0 + b), ep, kDomainRespawn, ((uint32_t)(tstep & 0xFFF) << 4) | (uint32_t)l);
      if (l == q) {
        tx = bits_to_pos<T>(w.x) * (T)0.95;
        ty = bits_to_pos<T>(w.y) * (T)0.95;
        moved = true;
      }
      nf = (nf & ~(1u << l)) | ((w.z >> 31) << l) | (1u << (6 + l));
    }
  }
  {  // deposit: a collector that (now) holds type h and touches deposit h lets go; each collector lane decides for itself
    const int h = coll ? (int)((nf >> (12 + 2 * q)) & 3u) - 1 : -1;
    const bool drop = h >= 0 && ((cdbits >> (h > 0 ? 1 : 0)) & 1u);
    if (__any_sync(0xffffffffu, drop)) nf &= ~__reduce_or_sync(env_mask, drop ? (3u << (12 + 2 * q)) : 0u);
  }
  T team = (T)0;
  if (s.track) {  // team return, added in agent order like the thread-per-env kernels
#pragma unroll
    for (int j = 0; j < kTrN; ++j) team += tr_shfl(r, base + j);
  }
  if (active) {
    if (moved) st2(lm_ptr, Vec2<T>{tx, ty});
    if (rew != nullptr) rew[b * kTrN + q] = r;
    if (done != nullptr) done[b * kTrN + q] = 0;  // no done_callback (experiments/scenarios.py:186-190): always False
    if (info_i != nullptr) info_i[b * (kTrN + 1) + q] = bench;
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
