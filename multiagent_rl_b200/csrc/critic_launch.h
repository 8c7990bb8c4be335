// Host-side interface of the critic forward kernel (critic_kernels.cu) used by cabi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpe {

constexpr int kCriticH = 64;        // dense1 width = LSTM input = LSTM hidden (ac_network_multi_gumbel.py:86-92)
constexpr int kCriticMaxIn = 64;    // obs_dim + sum(action widths)
constexpr int kCriticMaxAgents = 16;
constexpr int kCriticMaxOut = 4;

struct CriticHostWeights {
  const float *dense1_w, *dense1_b;          // [64][Din], [64]
  const float *w_ih, *w_hh, *b_ih, *b_hh;    // [256][64], [256][64], [256], [256]  (rows: i, f, g, o)
  const float *dense2_w, *dense2_b;          // [out][64], [out]
  const float *dense3_w, *dense3_b;          // [out][64], [out] or NULL
};

// Packed fp32 blob in device memory (what a CTA stages into shared memory with one bulk copy)
struct CriticDev {
  float *blob = nullptr;
  size_t blob_floats = 0;
  int32_t D = 0, A = 0, Din = 0, out = 1, has_r = 0, relu_attn = 1;
  int32_t off_w1 = 0, off_b1 = 0, off_wih = 0, off_whh = 0, off_bg = 0, off_w2 = 0, off_b2 = 0, off_w3 = 0, off_b3 = 0;
};

void critic_layout(int D, int A, int out, bool has_r, bool relu_attn, CriticDev *o);
void critic_pack(const CriticDev &d, const CriticHostWeights &w, float *host_blob);
// obs [B][N][D], action [B][N][A] (one-hot or soft), q / r [B][out]
cudaError_t launch_critic_forward(const CriticDev &w, const float *obs, const float *action, int64_t B, int N, float *q,
                                  float *r, cudaStream_t st);

}  // namespace mpe
