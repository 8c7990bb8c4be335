// Critic forward: cat(obs, action) -> relu(dense1) -> LSTM(64) over the agent axis -> dot-product attention with the
// final hidden state -> softmax over agents -> weighted sum (-> relu) -> dense2 (and dense3 for the "+model" critic).
//
// Reference rows: rls/model/ac_network_multi_gumbel.py:70-148 (CriticNetwork.forward :123-148, attention_net :94-121),
// rls/model/ac_network_model_multi_gumbel.py:69-143 (two heads, no relu after the attention), called from
// rls/agent/multiagent/ddpg_gumbel_fix.py:151,159,191 (optimize) and model_ddpg_gumbel_fix.py:155,163,199.
// not on the acting path; here so that imagined rollouts / TD targets can stay on the device.
//
// Mapping: one warp per sample.  A persistent CTA (8 warps) keeps the packed weights (W_ih, W_hh: 2 x 64 x 256 fp32 =
// 128 KB, plus dense1 / heads) in shared memory, loaded once with a TMA bulk copy.  Lane l owns hidden units l and
// l + 32: its 8 gate pre-activations per step are one 32 B slice of a k-major weight row, x_t / h_{t-1} are broadcast
// reads of the warp's staging row.  All arithmetic fp32 (FFMA), exp / tanh by the accurate libdevice forms.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "critic_launch.h"

namespace mpe {

constexpr int kCriticThreads = 256;
constexpr int kCriticWarps = kCriticThreads / 32;
constexpr int kGates = 4 * kCriticH;  // 256

static int round_up4(int x) { return (x + 3) / 4 * 4; }

void critic_layout(int D, int A, int out, bool has_r, bool relu_attn, CriticDev *o) {
  o->D = D; o->A = A; o->Din = D + A; o->out = out; o->has_r = has_r ? 1 : 0; o->relu_attn = relu_attn ? 1 : 0;
  int off = 0;
  o->off_wih = off; off += kCriticH * kGates;
  o->off_whh = off; off += kCriticH * kGates;
  o->off_bg = off; off += kGates;
  o->off_w1 = off; off += round_up4(o->Din * kCriticH);
  o->off_b1 = off; off += kCriticH;
  o->off_w2 = off; off += kCriticH * kCriticMaxOut;
  o->off_b2 = off; off += kCriticMaxOut;
  o->off_w3 = off; off += kCriticH * kCriticMaxOut;
  o->off_b3 = off; off += kCriticMaxOut;
  o->blob_floats = (size_t)((off + 31) / 32 * 32);
}

// reference gate row (gate * 64 + unit) -> packed column: lane = unit & 31 owns [i0 i1 f0 f1 g0 g1 o0 o1], units (lane, lane + 32)
static int packed_col(int gate, int unit) { return (unit & 31) * 8 + gate * 2 + (unit >> 5); }

void critic_pack(const CriticDev &d, const CriticHostWeights &w, float *blob) {
  std::memset(blob, 0, d.blob_floats * sizeof(float));
  for (int gate = 0; gate < 4; ++gate)
    for (int unit = 0; unit < kCriticH; ++unit) {
      const int row = gate * kCriticH + unit, col = packed_col(gate, unit);
      for (int k = 0; k < kCriticH; ++k) {
        blob[d.off_wih + k * kGates + col] = w.w_ih[row * kCriticH + k];
        blob[d.off_whh + k * kGates + col] = w.w_hh[row * kCriticH + k];
      }
      blob[d.off_bg + col] = w.b_ih[row] + w.b_hh[row];
    }
  for (int j = 0; j < kCriticH; ++j) {
    for (int k = 0; k < d.Din; ++k) blob[d.off_w1 + k * kCriticH + j] = w.dense1_w[j * d.Din + k];
    blob[d.off_b1 + j] = w.dense1_b[j];
  }
  for (int o = 0; o < d.out; ++o) {
    for (int k = 0; k < kCriticH; ++k) {
      blob[d.off_w2 + k * kCriticMaxOut + o] = w.dense2_w[o * kCriticH + k];
      if (d.has_r) blob[d.off_w3 + k * kCriticMaxOut + o] = w.dense3_w[o * kCriticH + k];
    }
    blob[d.off_b2 + o] = w.dense2_b[o];
    if (d.has_r) blob[d.off_b3 + o] = w.dense3_b[o];
  }
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-warp staging: the input row, x_t = relu(dense1), h_{t-1}, and every step's output for the attention
struct CriticWarpSmem {
  float in[kCriticMaxIn];
  float x[kCriticH];
  float h[kCriticH];
  float out[kCriticMaxAgents][kCriticH];
};

__global__ void __launch_bounds__(kCriticThreads, 1)
    k_critic_forward(CriticDev w, const float *__restrict__ obs, const float *__restrict__ action, int64_t B, int N,
                     float *__restrict__ q, float *__restrict__ r) {
  extern __shared__ __align__(128) float smem_f[];
  float *sw = smem_f;
  CriticWarpSmem *ws = reinterpret_cast<CriticWarpSmem *>(sw + w.blob_floats);
  uint64_t *bar = reinterpret_cast<uint64_t *>(ws + kCriticWarps);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (uint32_t)(w.blob_floats * sizeof(float)));
    bulk_load(sw, w.blob, (uint32_t)(w.blob_floats * sizeof(float)), bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  CriticWarpSmem &s = ws[warp];
  const float *Wih = sw + w.off_wih + lane * 8, *Whh = sw + w.off_whh + lane * 8, *bg = sw + w.off_bg + lane * 8;
  const float *W1 = sw + w.off_w1, *b1 = sw + w.off_b1;
  const int D = w.D, A = w.A, Din = w.Din;
  for (int64_t b = (int64_t)blockIdx.x * kCriticWarps + warp; b < B; b += (int64_t)gridDim.x * kCriticWarps) {
    float c0 = 0.0f, c1 = 0.0f, h0 = 0.0f, h1 = 0.0f;
    s.h[lane] = 0.0f; s.h[lane + 32] = 0.0f;
    for (int t = 0; t < N; ++t) {
      // obs_act = cat(obs, action) (ac_network_multi_gumbel.py:131-134)
      __syncwarp();
      for (int k = lane; k < Din; k += 32)
        s.in[k] = k < D ? obs[(b * N + t) * D + k] : action[(b * N + t) * A + (k - D)];
      __syncwarp();
      // relu(dense1): lane computes outputs lane and lane + 32
      float a0 = b1[lane], a1 = b1[lane + 32];
      for (int k = 0; k < Din; ++k) {
        const float xk = s.in[k];
        a0 = fmaf(xk, W1[k * kCriticH + lane], a0);
        a1 = fmaf(xk, W1[k * kCriticH + lane + 32], a1);
      }
      s.x[lane] = fmaxf(a0, 0.0f); s.x[lane + 32] = fmaxf(a1, 0.0f);
      __syncwarp();
      // gates = b + W_ih x_t + W_hh h_{t-1}; packed columns [i0 i1 f0 f1 g0 g1 o0 o1] of units (lane, lane + 32)
      float acc[8];
      {
        const float4 u = *reinterpret_cast<const float4 *>(bg), v = *reinterpret_cast<const float4 *>(bg + 4);
        acc[0] = u.x; acc[1] = u.y; acc[2] = u.z; acc[3] = u.w; acc[4] = v.x; acc[5] = v.y; acc[6] = v.z; acc[7] = v.w;
      }
#pragma unroll 4
      for (int k = 0; k < kCriticH; ++k) {
        const float xk = s.x[k], hk = s.h[k];
        const float4 u = *reinterpret_cast<const float4 *>(Wih + k * kGates), v = *reinterpret_cast<const float4 *>(Wih + k * kGates + 4);
        const float4 p = *reinterpret_cast<const float4 *>(Whh + k * kGates), z = *reinterpret_cast<const float4 *>(Whh + k * kGates + 4);
        acc[0] = fmaf(xk, u.x, acc[0]); acc[1] = fmaf(xk, u.y, acc[1]); acc[2] = fmaf(xk, u.z, acc[2]); acc[3] = fmaf(xk, u.w, acc[3]);
        acc[4] = fmaf(xk, v.x, acc[4]); acc[5] = fmaf(xk, v.y, acc[5]); acc[6] = fmaf(xk, v.z, acc[6]); acc[7] = fmaf(xk, v.w, acc[7]);
        acc[0] = fmaf(hk, p.x, acc[0]); acc[1] = fmaf(hk, p.y, acc[1]); acc[2] = fmaf(hk, p.z, acc[2]); acc[3] = fmaf(hk, p.w, acc[3]);
        acc[4] = fmaf(hk, z.x, acc[4]); acc[5] = fmaf(hk, z.y, acc[5]); acc[6] = fmaf(hk, z.z, acc[6]); acc[7] = fmaf(hk, z.w, acc[7]);
      }
      c0 = sigmoid_acc(acc[2]) * c0 + sigmoid_acc(acc[0]) * tanhf(acc[4]);
      c1 = sigmoid_acc(acc[3]) * c1 + sigmoid_acc(acc[1]) * tanhf(acc[5]);
      h0 = sigmoid_acc(acc[6]) * tanhf(c0);
      h1 = sigmoid_acc(acc[7]) * tanhf(c1);
      __syncwarp();  // every lane has read h_{t-1}
      s.h[lane] = h0; s.h[lane + 32] = h1;
      s.out[t][lane] = h0; s.out[t][lane + 32] = h1;
    }
    __syncwarp();
    // attention_net (:103-110): scores = <output_t, h_N>, softmax over the agent axis, weighted sum of the outputs
    float score[kCriticMaxAgents];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kCriticMaxAgents; ++t) {
      if (t < N) {
        score[t] = warp_sum(s.out[t][lane] * h0 + s.out[t][lane + 32] * h1);
        mx = fmaxf(mx, score[t]);
      }
    }
    float den = 0.0f;
#pragma unroll
    for (int t = 0; t < kCriticMaxAgents; ++t)
      if (t < N) { score[t] = expf(score[t] - mx); den += score[t]; }
    float n0 = 0.0f, n1 = 0.0f;
#pragma unroll
    for (int t = 0; t < kCriticMaxAgents; ++t)
      if (t < N) {
        const float a = score[t] / den;
        n0 = fmaf(s.out[t][lane], a, n0);
        n1 = fmaf(s.out[t][lane + 32], a, n1);
      }
    if (w.relu_attn) { n0 = fmaxf(n0, 0.0f); n1 = fmaxf(n1, 0.0f); }  // ac_network_multi_gumbel.py:141 only
    for (int o = 0; o < w.out; ++o) {
      const float *W2 = sw + w.off_w2;
      const float v = warp_sum(n0 * W2[lane * kCriticMaxOut + o] + n1 * W2[(lane + 32) * kCriticMaxOut + o]);
      if (lane == 0) q[b * w.out + o] = v + sw[w.off_b2 + o];
      if (w.has_r && r != nullptr) {
        const float *W3 = sw + w.off_w3;
        const float u = warp_sum(n0 * W3[lane * kCriticMaxOut + o] + n1 * W3[(lane + 32) * kCriticMaxOut + o]);
        if (lane == 0) r[b * w.out + o] = u + sw[w.off_b3 + o];
      }
    }
  }
}

cudaError_t launch_critic_forward(const CriticDev &w, const float *obs, const float *action, int64_t B, int N, float *q,
                                  float *r, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (N < 1 || N > kCriticMaxAgents) return cudaErrorInvalidValue;
  const size_t smem = w.blob_floats * sizeof(float) + kCriticWarps * sizeof(CriticWarpSmem) + 64;
  static size_t have[64] = {0};
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (have[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(k_critic_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    have[dev] = smem;
  }
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (B + kCriticWarps - 1) / kCriticWarps;
  const int grid = (int)(want < nsm ? want : nsm);
  k_critic_forward<<<grid, kCriticThreads, smem, st>>>(w, obs, action, B, N, q, r);
  return cudaGetLastError();
}

}  // namespace mpe
