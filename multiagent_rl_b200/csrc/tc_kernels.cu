// Tensor-core (tcgen05 / TMEM) version of the actor forward and of the fused rollout step, N <= 3 agents.
//
// Reference rows: same as actor_kernels.cu (rls/model/ac_network_multi_gumbel.py:52-67,
// rls/agent/multiagent/ddpg_gumbel_fix.py:86-116, experiments/run.py:36-65).
//
// Why tensor cores: ncu on the fp32 SIMT kernel shows DRAM 0.2 %, FMA pipe 54 %, issue 64 % - the batched
// LSTM GEMMs are contraction-bound.  Why fp16 hi/lo: the acting path must reproduce the reference's fp32
// logits to ~1e-6 so that the sampled index is bit-exact under identical noise; every fp32 operand x is
// split as x = hi + lo (two fp16 values, 22 significant bits) and each product is three kind::f16 MMAs
// (hi*hi + lo*hi + hi*lo) accumulated in fp32 in TMEM.  Weights are pre-split (and pre-scaled by 16 so
// that the lo halves stay normal) on the host; activations are split by the epilogue that produces them.
//
// One persistent CTA per SM, one tile = 128 envs = 128 TMEM lanes, thread r of each warpgroup owns env row r:
//   warps 0-3  (WG0): epilogue of the forward  LSTM direction, sampling, physics/reward of env r
//   warps 4-7  (WG1): epilogue of the backward LSTM direction
//   warp  8         : lane 0 issues every tcgen05.mma of the CTA (one ordered stream)
// TMEM columns:  [0,256)  gate accumulators G[dir] (aliased by the dense1 accumulators D1[t] before the LSTM)
//                [256, 256+64N)  h1[t] as an fp16 hi/lo A operand (never leaves TMEM)
//                [256+64N, +16N) logits[t] accumulators
// Shared memory: fp16 hi/lo weight image (one TMA bulk load), obs A operands, recurrent h / relu(h) A operands.
// The two directions form a natural ping-pong: while WG0 runs the cell math of step s, the tensor pipe runs
// the gate GEMM of the other direction.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "actor_launch.h"
#include "env_core.cuh"
#include "tc_common.cuh"

namespace mpe {

constexpr int kTcThreads = 288;
constexpr int kRows = 128;
constexpr float kWScale = 16.0f, kWInv = 0.0625f;
constexpr uint32_t kChunkA = kRows * 16;  // bytes of one K-chunk (8 halves) of a 128-row A operand

// ------------------------------------------------------------------------------------------------
// host: fp16 hi/lo weight image
// ------------------------------------------------------------------------------------------------
void tc_layout(int D, int A0, int A1, TcDev *o) {
  o->D = D; o->A0 = A0; o->A1 = A1; o->A = A0 + A1;
  o->Kx = D <= 16 ? 16 : (D <= 32 ? 32 : 0);
  uint32_t off = 0;
  for (int d = 0; d < 2; ++d)
    for (int hl = 0; hl < 2; ++hl) { o->off_wih[d][hl] = off; off += kHid * kGateN * 2; }
  for (int d = 0; d < 2; ++d)
    for (int hl = 0; hl < 2; ++hl) { o->off_whh[d][hl] = off; off += kH * kGateN * 2; }
  for (int hl = 0; hl < 2; ++hl) { o->off_w1[hl] = off; off += (o->Kx > 0 ? o->Kx : 16) * kHid * 2; }
  for (int d = 0; d < 2; ++d)
    for (int hl = 0; hl < 2; ++hl) { o->off_w2[d][hl] = off; off += kH * 16 * 2; }
  o->off_bg = off; off += 2 * kGateN * 4;
  o->off_b1 = off; off += kHid * 4;
  o->off_b2 = off; off += 16 * 4;
  o->bytes = (off + 127) / 128 * 128;
}

bool tc_supported(const TcDev &t) { return t.Kx > 0 && t.A <= 16; }

static void put_split(unsigned char *img, uint32_t off_hi, uint32_t off_lo, int rows, int n, int k, float v) {
  const float x = v * kWScale;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  const size_t idx = ((size_t)(k / 8) * rows + n) * 8 + (k % 8);
  reinterpret_cast<__half *>(img + off_hi)[idx] = hi;
  reinterpret_cast<__half *>(img + off_lo)[idx] = lo;
}

void tc_pack(const TcDev &t, const ActorHostWeights &w, unsigned char *img) {
  std::memset(img, 0, t.bytes);
  const float *wih[2] = {w.w_ih, w.w_ih_r}, *whh[2] = {w.w_hh, w.w_hh_r};
  const float *bih[2] = {w.b_ih, w.b_ih_r}, *bhh[2] = {w.b_hh, w.b_hh_r};
  float *bg = reinterpret_cast<float *>(img + t.off_bg);
  for (int d = 0; d < 2; ++d)
    for (int g = 0; g < 4; ++g)
      for (int u = 0; u < kH; ++u) {
        const int row = g * kH + u, n = u * 4 + g;  // packed gate column: unit-major, [i f g o] adjacent
        for (int k = 0; k < kHid; ++k) put_split(img, t.off_wih[d][0], t.off_wih[d][1], kGateN, n, k, wih[d][row * kHid + k]);
        for (int k = 0; k < kH; ++k) put_split(img, t.off_whh[d][0], t.off_whh[d][1], kGateN, n, k, whh[d][row * kH + k]);
        bg[d * kGateN + n] = bih[d][row] + bhh[d][row];
      }
  float *b1 = reinterpret_cast<float *>(img + t.off_b1), *b2 = reinterpret_cast<float *>(img + t.off_b2);
  for (int j = 0; j < kHid; ++j) {
    for (int k = 0; k < t.D; ++k) put_split(img, t.off_w1[0], t.off_w1[1], kHid, j, k, w.dense1_w[j * t.D + k]);
    b1[j] = w.dense1_b[j];
  }
  for (int a = 0; a < t.A; ++a) {
    const float *src = a < t.A0 ? w.dense2_w + a * kHid : w.dense2b_w + (a - t.A0) * kHid;
    for (int d = 0; d < 2; ++d)
      for (int u = 0; u < kH; ++u) put_split(img, t.off_w2[d][0], t.off_w2[d][1], 16, a, u, src[d * kH + u]);
    b2[a] = a < t.A0 ? w.dense2_b[a] : w.dense2b_b[a - t.A0];
  }
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_tc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_tc(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 8 fp32 values -> one 16 B vector of fp16 hi parts and one of lo parts, stored at the row's slot of a K-chunk
__device__ __forceinline__ void store_chunk_split(unsigned char *hi_chunk, unsigned char *lo_chunk, int row,
                                                  const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half h0, l0, h1, l1;
    split_f16(v[2 * i], h0, l0);
    split_f16(v[2 * i + 1], h1, l1);
    h[i] = pack_h2(h0, h1);
    l[i] = pack_h2(l0, l1);
  }
  *reinterpret_cast<uint4 *>(hi_chunk + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(lo_chunk + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

// three-term split product D (+)= A*B with A and B in shared memory, K = 16 * kblocks
__device__ __forceinline__ void mma3_ss(uint32_t tmem_d, const unsigned char *a_hi, const unsigned char *a_lo,
                                        uint32_t a_lbo, const unsigned char *b_hi, const unsigned char *b_lo,
                                        uint32_t b_lbo, int kblocks, uint32_t idesc, bool accumulate) {
  for (int kb = 0; kb < kblocks; ++kb) {
    const uint64_t ah = make_smem_desc(a_hi + kb * 2 * a_lbo, a_lbo, 128), al = make_smem_desc(a_lo + kb * 2 * a_lbo, a_lbo, 128);
    const uint64_t bh = make_smem_desc(b_hi + kb * 2 * b_lbo, b_lbo, 128), bl = make_smem_desc(b_lo + kb * 2 * b_lbo, b_lbo, 128);
    mma_f16_ss(tmem_d, ah, bh, idesc, accumulate || kb > 0);
    mma_f16_ss(tmem_d, al, bh, idesc, true);
    mma_f16_ss(tmem_d, ah, bl, idesc, true);
  }
}
// same with the A operand in TMEM (hi at column a_hi, lo at column a_lo; 8 columns per K = 16 block)
__device__ __forceinline__ void mma3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, const unsigned char *b_hi,
                                        const unsigned char *b_lo, uint32_t b_lbo, int kblocks, uint32_t idesc,
                                        bool accumulate) {
  for (int kb = 0; kb < kblocks; ++kb) {
    const uint64_t bh = make_smem_desc(b_hi + kb * 2 * b_lbo, b_lbo, 128), bl = make_smem_desc(b_lo + kb * 2 * b_lbo, b_lbo, 128);
    mma_f16_ts(tmem_d, a_hi + kb * 8, bh, idesc, accumulate || kb > 0);
    mma_f16_ts(tmem_d, a_lo + kb * 8, bh, idesc, true);
    mma_f16_ts(tmem_d, a_hi + kb * 8, bl, idesc, true);
  }
}

// Optional phase timeline (MPE_TC_TIMELINE=1): clock64 stamps of the first tile of every CTA, read back with
// the (undeclared, debug-only) export mpe_debug_tc_timeline.
__device__ unsigned long long g_tc_timeline[148 * 96];
#define TL(role, i)                                                                         \
  do {                                                                                      \
    if (dbg && first_tile && blockIdx.x < 148) g_tc_timeline[blockIdx.x * 96 + (role) * 32 + (i)] = clock64(); \
  } while (0)

struct TcSmem {
  unsigned char *w, *x, *h, *r;  // weight image; obs operands [N][hl][Kx/8][128][8]; h and relu(h) [dir][hl][4][128][8]
  float *stage_obs, *stage_rew;  // fp32 row staging for TMA loads/stores (aliases h / r, free outside the LSTM)
  int *act;                      // [128][N][2]
  uint64_t *bars;                // see enum below
  uint32_t *tmem_slot;
};
enum { B_W = 0, B_X, B_D1, B_H1, B_G0, B_G1, B_H0, B_H1R, B_L, B_OBS, B_COUNT };

__host__ __device__ inline size_t tc_smem_bytes(uint32_t wbytes, int N, int Kx) {
  return (size_t)wbytes + (size_t)N * 2 * (Kx / 8) * kChunkA + 2 * 32768 + (size_t)kRows * N * 2 * 4 + B_COUNT * 8 + 64;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int SC, int N, bool FUSED>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_tc(EnvState<float> s, TcDev w, ActorIO io, RolloutIO ro, int max_episode_len, int64_t ntiles, int dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  using Dm = Dims<SC, N>;
  const int D = FUSED ? Dm::D : w.D;
  const int R = N * D;
  const int Kx = w.Kx;
  const int tid = threadIdx.x, warp = tid >> 5;
  TcSmem sm;
  sm.w = smem;
  sm.x = smem + w.bytes;
  sm.h = sm.x + (size_t)N * 2 * (Kx / 8) * kChunkA;
  sm.r = sm.h + 32768;
  sm.stage_obs = reinterpret_cast<float *>(sm.h);
  sm.stage_rew = reinterpret_cast<float *>(sm.r);
  sm.act = reinterpret_cast<int *>(sm.r + 32768);
  sm.bars = reinterpret_cast<uint64_t *>(sm.act + kRows * N * 2);
  sm.tmem_slot = reinterpret_cast<uint32_t *>(sm.bars + B_COUNT);

  if (tid == 0) {
    mbar_init(&sm.bars[B_W], 1);
    mbar_init(&sm.bars[B_X], 256);
    mbar_init(&sm.bars[B_D1], 1);
    mbar_init(&sm.bars[B_H1], 256);
    mbar_init(&sm.bars[B_G0], 1);
    mbar_init(&sm.bars[B_G1], 1);
    mbar_init(&sm.bars[B_H0], 128);
    mbar_init(&sm.bars[B_H1R], 128);
    mbar_init(&sm.bars[B_L], 1);
    mbar_init(&sm.bars[B_OBS], 1);
    mbar_fence_init();
    mbar_expect_tx(&sm.bars[B_W], w.bytes);
    bulk_load(sm.w, w.blob, w.bytes, &sm.bars[B_W]);
  }
  if (warp == 8) tmem_alloc(sm.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem_slot;
  const uint32_t col_h1 = 256, col_l = 256 + 64 * N;
  const int T = FUSED ? ro.T : 1;

  if (warp == 8) {
    // =============================== MMA issuer ===============================
    if ((tid & 31) == 0) {
      uint32_t ph_x = 0, ph_h1 = 0, ph_h[2] = {0, 0};
      const uint32_t id_g = make_idesc_f16(128, 128), id_d1 = make_idesc_f16(128, 64), id_l = make_idesc_f16(128, 16);
      mbar_wait(&sm.bars[B_W], 0);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const bool first_tile = tile == blockIdx.x;
        for (int it = 0; it < T; ++it) {
          mbar_wait(&sm.bars[B_X], ph_x); ph_x ^= 1;
          tc_fence_after();
          TL(2, 0);
          for (int t = 0; t < N; ++t) {  // dense1: D1[t] = x[t] * W1^T
            const unsigned char *xh = sm.x + (size_t)(t * 2) * (Kx / 8) * kChunkA, *xl = xh + (size_t)(Kx / 8) * kChunkA;
            mma3_ss(tmem + t * 64, xh, xl, kChunkA, sm.w + w.off_w1[0], sm.w + w.off_w1[1], kHid * 16, Kx / 16, id_d1, false);
          }
          mma_commit(&sm.bars[B_D1]);
          TL(2, 1);
          mbar_wait(&sm.bars[B_H1], ph_h1); ph_h1 ^= 1;
          tc_fence_after();
          TL(2, 2);
          uint32_t seen = 0;
          for (int st = 0; st <= N; ++st) {
            for (int d = 0; d < 2; ++d) {
              if (st > 0) {  // h of step st-1 is in smem: its dense2 contribution, then (if any) the next gates
                mbar_wait(&sm.bars[B_H0 + d], ph_h[d]); ph_h[d] ^= 1;
                tc_fence_after();
                TL(2, 3 + st * 4 + d * 2);
                const int tp = d == 0 ? st - 1 : N - st;
                const unsigned char *rh = sm.r + d * 16384, *rl = rh + 8192;
                mma3_ss(tmem + col_l + tp * 16, rh, rl, kChunkA, sm.w + w.off_w2[d][0], sm.w + w.off_w2[d][1], 16 * 16, 2,
                        id_l, (seen >> tp) & 1);
                seen |= 1u << tp;
              }
              if (st < N) {
                const int t = d == 0 ? st : N - 1 - st;
                mma3_ts(tmem + d * 128, tmem + col_h1 + t * 64, tmem + col_h1 + t * 64 + 32, sm.w + w.off_wih[d][0],
                        sm.w + w.off_wih[d][1], kGateN * 16, 4, id_g, false);
                if (st > 0) {
                  const unsigned char *hh = sm.h + d * 16384, *hl = hh + 8192;
                  mma3_ss(tmem + d * 128, hh, hl, kChunkA, sm.w + w.off_whh[d][0], sm.w + w.off_whh[d][1], kGateN * 16, 2,
                          id_g, true);
                }
                mma_commit(&sm.bars[B_G0 + d]);
                TL(2, 4 + st * 4 + d * 2);
              }
            }
          }
          mma_commit(&sm.bars[B_L]);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue warpgroups ===============================
    const int d = tid >> 7, row = tid & 127;
    const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
    const float *bg = reinterpret_cast<const float *>(sm.w + w.off_bg) + d * kGateN;
    const float *b1 = reinterpret_cast<const float *>(sm.w + w.off_b1);
    const float *b2 = reinterpret_cast<const float *>(sm.w + w.off_b2);
    uint32_t ph_d1 = 0, ph_g = 0, ph_l = 0, ph_obs = 0;
    mbar_wait(&sm.bars[B_W], 0);  // biases are read from the weight image
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t env0 = tile * kRows;
      const int64_t nb = FUSED ? s.B : io.B;
      const int valid = (int)((nb - env0) < kRows ? (nb - env0) : kRows);
      const bool mine = row < valid;
      const int64_t b = env0 + row;
      const bool first_tile = tile == blockIdx.x && row == 0;
      TL(d, 0);
      if (!FUSED) {  // obs tile -> fp32 staging (TMA when whole and aligned)
        const float *src = io.obs + env0 * R;
        if (valid == kRows && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          if (tid == 0) {
            mbar_expect_tx(&sm.bars[B_OBS], (uint32_t)(kRows * R * 4));
            bulk_load(sm.stage_obs, src, (uint32_t)(kRows * R * 4), &sm.bars[B_OBS]);
          }
          mbar_wait(&sm.bars[B_OBS], ph_obs); ph_obs ^= 1;
        } else {
          for (int i = tid; i < kRows * R; i += 256) sm.stage_obs[i] = i < valid * R ? src[i] : 0.0f;
          bar_sync_n(1, 256);
        }
      }
      for (int it = 0; it < T; ++it) {
        // ---- observations -> fp16 hi/lo A operands (agents split between the warpgroups) ----
        {
          Env<float, SC, N> e;
          float comm[2][10];
          if (FUSED) {
            if (mine) {
              e.load(s, b);
              if (SC == kReference) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                  for (int k = 0; k < 10; ++k) comm[i][k] = s.comm[((int64_t)i * 10 + k) * s.B + b];
              }
            }
          }
#pragma unroll
          for (int t = 0; t < N; ++t) {
            if ((t & 1) != d) continue;
            float xr[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) xr[k] = 0.0f;
            if (FUSED) {
              if (mine) e.obs_row(t, xr, SC == kReference ? comm[1 - (t & 1)] : nullptr);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k < D) xr[k] = sm.stage_obs[row * R + t * D + k];
            }
            unsigned char *xh = sm.x + (size_t)(t * 2) * (Kx / 8) * kChunkA, *xl = xh + (size_t)(Kx / 8) * kChunkA;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c * 8 < Kx) {
                const float v[8] = {xr[c * 8], xr[c * 8 + 1], xr[c * 8 + 2], xr[c * 8 + 3],
                                    xr[c * 8 + 4], xr[c * 8 + 5], xr[c * 8 + 6], xr[c * 8 + 7]};
                store_chunk_split(xh + c * kChunkA, xl + c * kChunkA, row, v);
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&sm.bars[B_X]);
        TL(d, 1);

        // ---- dense1 epilogue: h1 = relu(D1/16 + b1) -> fp16 hi/lo A operand in TMEM ----
        mbar_wait(&sm.bars[B_D1], ph_d1); ph_d1 ^= 1;
        tc_fence_after();
        TL(d, 2);
#pragma unroll 1
        for (int t = 0; t < N; ++t) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + t * 64 + d * 32, v);
          tmem_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a0 = fmaxf(fmaf(__uint_as_float(v[2 * j]), kWInv, b1[d * 32 + 2 * j]), 0.0f);
            const float a1 = fmaxf(fmaf(__uint_as_float(v[2 * j + 1]), kWInv, b1[d * 32 + 2 * j + 1]), 0.0f);
            __half h0, l0, h1, l1;
            split_f16(a0, h0, l0);
            split_f16(a1, h1, l1);
            hi[j] = pack_h2(h0, h1);
            lo[j] = pack_h2(l0, l1);
          }
          tmem_st16(tmem + lane_base + col_h1 + t * 64 + d * 16, hi);
          tmem_st16(tmem + lane_base + col_h1 + t * 64 + 32 + d * 16, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sm.bars[B_H1]);
        TL(d, 3);

        // ---- LSTM cell math of this warpgroup's direction ----
        float c[kH];
#pragma unroll
        for (int u = 0; u < kH; ++u) c[u] = 0.0f;
        unsigned char *hh = sm.h + d * 16384, *hl = hh + 8192, *rh = sm.r + d * 16384, *rl = rh + 8192;
#pragma unroll 1
        for (int st = 0; st < N; ++st) {
          mbar_wait(&sm.bars[B_G0 + d], ph_g); ph_g ^= 1;
          tc_fence_after();
          TL(d, 4 + 2 * st);
#pragma unroll
          for (int ub = 0; ub < 4; ++ub) {  // 8 units x [i f g o] = 32 accumulator columns
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + d * 128 + ub * 32, v);
            tmem_wait_ld();
            float hv[8], rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = *reinterpret_cast<const float4 *>(bg + ub * 32 + j * 4);
              const float ig = sigmoid_tc(fmaf(__uint_as_float(v[4 * j]), kWInv, bb.x));
              const float fg = sigmoid_tc(fmaf(__uint_as_float(v[4 * j + 1]), kWInv, bb.y));
              const float gg = tanh_tc(fmaf(__uint_as_float(v[4 * j + 2]), kWInv, bb.z));
              const float og = sigmoid_tc(fmaf(__uint_as_float(v[4 * j + 3]), kWInv, bb.w));
              const float cn = fmaf(fg, c[ub * 8 + j], ig * gg);
              c[ub * 8 + j] = cn;
              hv[j] = og * tanh_tc(cn);
              rv[j] = fmaxf(hv[j], 0.0f);
            }
            store_chunk_split(hh + ub * kChunkA, hl + ub * kChunkA, row, hv);
            store_chunk_split(rh + ub * kChunkA, rl + ub * kChunkA, row, rv);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&sm.bars[B_H0 + d]);
          TL(d, 5 + 2 * st);
        }

        // ---- heads: logits -> Gumbel-max sample -> (fused) physics, reward, outputs ----
        if (d == 0) {
          mbar_wait(&sm.bars[B_L], ph_l); ph_l ^= 1;
          tc_fence_after();
          TL(d, 12);
          int au[N], ac[N];
          const uint64_t step = FUSED ? ro.step0 + (uint64_t)it : io.step;
          const uint64_t seed = FUSED ? s.seed : io.seed;
          const int64_t gid0 = FUSED ? s.gid0 : io.gid0;
#pragma unroll
          for (int t = 0; t < N; ++t) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + col_l + t * 16, v);
            tmem_wait_ld();
            float lg[16], z[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) lg[a] = fmaf(__uint_as_float(v[a]), kWInv, b2[a]);
            const int64_t orow = b * N + t;
            if (!FUSED && io.gumbel != nullptr) {
#pragma unroll
              for (int a = 0; a < 16; ++a) z[a] = (a < w.A && mine) ? lg[a] + io.gumbel[orow * w.A + a] : lg[a];
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (4 * j < w.A) {
                  const uint4 rr = philox_raw(seed, (uint64_t)(gid0 + b), (uint32_t)step, kDomainGumbel, t * 8 + j);
                  z[4 * j] = lg[4 * j] + bits_to_gumbel(rr.x); z[4 * j + 1] = lg[4 * j + 1] + bits_to_gumbel(rr.y);
                  z[4 * j + 2] = lg[4 * j + 2] + bits_to_gumbel(rr.z); z[4 * j + 3] = lg[4 * j + 3] + bits_to_gumbel(rr.w);
                } else {
                  z[4 * j] = z[4 * j + 1] = z[4 * j + 2] = z[4 * j + 3] = 0.0f;
                }
              }
            }
            int bu = 0, bc = 0;
            float best = z[0];
#pragma unroll
            for (int a = 1; a < 16; ++a)
              if (a < w.A0 && z[a] > best) { best = z[a]; bu = a; }
            if (w.A1 > 0) {
              float bcv = -INFINITY;
#pragma unroll
              for (int a = 0; a < 16; ++a)
                if (a >= w.A0 && a < w.A && z[a] > bcv) { bcv = z[a]; bc = a - w.A0; }
            }
            au[t] = bu; ac[t] = bc;
            if (t == N - 1) TL(d, 15);
            sm.act[(row * N + t) * 2] = bu;
            sm.act[(row * N + t) * 2 + 1] = bc;
            if (!FUSED && mine && io.logits != nullptr) {
#pragma unroll
              for (int a = 0; a < 16; ++a)
                if (a < w.A) io.logits[orow * w.A + a] = lg[a];
            }
          }
          if (FUSED) {
            const int64_t toff = (int64_t)it * s.B;
            bool do_reset = false;
            Env<float, SC, N> e;
            float comm[2][10];
            double ret = 0.0, n_ep = 0.0, n_steps = 0.0;
            if (mine) {
              e.load(s, b);
              e.physics(au, s.max_speed, s.accel);
              if (SC == kReference) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                  for (int k = 0; k < 10; ++k) comm[i][k] = k == ac[i] ? 1.0f : 0.0f;
              }
              float r[N];
              int coll[N], occ;
              float md;
              e.reward(r, coll, occ, md);
              float sum = 0.0f;
#pragma unroll
              for (int i = 0; i < N; ++i) { sum += r[i]; sm.stage_rew[row * N + i] = r[i]; }
              const float ep_ret = s.ep_ret[b] + sum;
              const int ts = s.tstep[b] + 1;
              do_reset = max_episode_len > 0 && ts >= max_episode_len;
#pragma unroll
              for (int i = 0; i < N; ++i)
                e.obs_row(i, sm.stage_obs + row * R + i * D, SC == kReference ? comm[1 - (i & 1)] : nullptr);
              if (do_reset) {
                ret = (double)ep_ret; n_ep = 1.0; n_steps = (double)ts;
                s.ep_ret[b] = 0.0f; s.tstep[b] = 0;
              } else {
                s.ep_ret[b] = ep_ret; s.tstep[b] = ts;
              }
            }
            fold_stats(s.stats, ret, n_ep, n_steps);
            float *g_obs = ro.obs_next != nullptr ? ro.obs_next + (toff + env0) * R : nullptr;
            float *g_rew = ro.rew != nullptr ? ro.rew + (toff + env0) * N : nullptr;
            const bool tma_ok = valid == kRows && ((reinterpret_cast<uintptr_t>(g_obs) | reinterpret_cast<uintptr_t>(g_rew)) & 15) == 0;
            if (g_obs != nullptr || g_rew != nullptr) {
              if (tma_ok) {
                fence_proxy_async_smem();
                bar_sync_n(2, 128);
                if (tid == 0) {
                  if (g_obs != nullptr) bulk_store(g_obs, sm.stage_obs, kRows * R * 4);
                  if (g_rew != nullptr) bulk_store(g_rew, sm.stage_rew, kRows * N * 4);
                  bulk_commit();
                  bulk_wait_read_all();
                }
              } else {
                bar_sync_n(2, 128);
                if (g_obs != nullptr)
                  for (int i = row; i < valid * R; i += 128) g_obs[i] = sm.stage_obs[i];
                if (g_rew != nullptr)
                  for (int i = row; i < valid * N; i += 128) g_rew[i] = sm.stage_rew[i];
              }
            }
            bar_sync_n(2, 128);
            if (ro.act_u != nullptr)
              for (int i = row; i < valid * N; i += 128) ro.act_u[(toff + env0) * N + i] = sm.act[i * 2];
            if (ro.act_c != nullptr)
              for (int i = row; i < valid * N; i += 128) ro.act_c[(toff + env0) * N + i] = sm.act[i * 2 + 1];
            if (mine) {
              if (do_reset) {
                const uint32_t ep = s.episode[b] + 1u;
                s.episode[b] = ep;
                e.reset(s.seed, (uint64_t)(s.gid0 + b), ep);
                e.store_world(s, b);
                if (SC == kReference) {
#pragma unroll
                  for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int k = 0; k < 10; ++k) comm[i][k] = 0.0f;
                }
              }
              e.store_agents(s, b);
              if (SC == kReference) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                  for (int k = 0; k < 10; ++k) s.comm[((int64_t)i * 10 + k) * s.B + b] = comm[i][k];
              }
            }
            __threadfence_block();
          } else {
            bar_sync_n(2, 128);
            const int rows = valid * N;
            if (io.act_u != nullptr)
              for (int i = row; i < rows; i += 128) io.act_u[env0 * N + i] = sm.act[i * 2];
            if (io.act_c != nullptr)
              for (int i = row; i < rows; i += 128) io.act_c[env0 * N + i] = sm.act[i * 2 + 1];
            if (io.onehot != nullptr)
              for (int i = row; i < rows * w.A; i += 128) {
                const int rr = i / w.A, a = i - rr * w.A;
                const bool hot = a < w.A0 ? (a == sm.act[rr * 2]) : (a - w.A0 == sm.act[rr * 2 + 1]);
                io.onehot[env0 * N * w.A + i] = hot ? 1.0f : 0.0f;
              }
          }
        }
        TL(d, 13);
        bar_sync_n(1, 256);  // both warpgroups: the state / staging of this iteration is settled
        TL(d, 14);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_free(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
static int sm_count_tc() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

template <int SC, int N, bool FUSED>
static cudaError_t launch_tc_t(const EnvState<float> &s, const TcDev &w, const ActorIO &io, const RolloutIO &ro,
                               int max_episode_len, int64_t nenvs, cudaStream_t st) {
  const size_t smem = tc_smem_bytes(w.bytes, N, w.Kx);
  cudaError_t e = cudaFuncSetAttribute(k_tc<SC, N, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (nenvs + kRows - 1) / kRows;
  const int nsm = sm_count_tc();
  const int grid = (int)(ntiles < nsm ? ntiles : nsm);
  static const int dbg = getenv("MPE_TC_TIMELINE") != nullptr;
  k_tc<SC, N, FUSED><<<grid, kTcThreads, smem, st>>>(s, w, io, ro, max_episode_len, ntiles, dbg);
  return cudaGetLastError();
}

bool tc_actor_supported(const TcDev &w, int N) { return tc_supported(w) && (N == 2 || N == 3); }

cudaError_t launch_actor_forward_tc(const TcDev &w, const ActorIO &io, cudaStream_t st) {
  EnvState<float> s{};
  RolloutIO ro;
  switch (io.N) {
    case 2: return launch_tc_t<kSpread, 2, false>(s, w, io, ro, 0, io.B, st);
    case 3: return launch_tc_t<kSpread, 3, false>(s, w, io, ro, 0, io.B, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_rollout_tc(const EnvStateAny &a, const TcDev &w, const RolloutIO &ro, cudaStream_t st) {
  EnvState<float> s;
  s.pv = static_cast<float *>(a.pv); s.lm = static_cast<float *>(a.lm); s.goal = a.goal; s.episode = a.episode;
  s.tstep = a.tstep; s.ep_ret = static_cast<float *>(a.ep_ret); s.comm = static_cast<float *>(a.comm);
  s.stats = a.stats; s.B = a.B; s.gid0 = a.gid0; s.seed = a.seed; s.max_speed = (float)a.max_speed;
  s.accel = (float)a.accel; s.track = 1;
  ActorIO io;
  if (a.scenario == kReference) return launch_tc_t<kReference, 2, true>(s, w, io, ro, a.max_episode_len, a.B, st);
  if (a.scenario == kSpeaker) return launch_tc_t<kSpeaker, 2, true>(s, w, io, ro, a.max_episode_len, a.B, st);
  if (a.N == 2) return launch_tc_t<kSpread, 2, true>(s, w, io, ro, a.max_episode_len, a.B, st);
  if (a.N == 3) return launch_tc_t<kSpread, 3, true>(s, w, io, ro, a.max_episode_len, a.B, st);
  return cudaErrorInvalidValue;
}

}  // namespace mpe

// debug-only export (not part of include/mpe_b200.h)
extern "C" __attribute__((visibility("default"))) int mpe_debug_tc_timeline(unsigned long long *out, int n) {
  return (int)cudaMemcpyFromSymbol(out, mpe::g_tc_timeline, sizeof(unsigned long long) * (size_t)n);
}
