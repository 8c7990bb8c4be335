// Host-side interface of the actor / fused-rollout kernels (actor_kernels.cu) used by cabi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_launch.h"

namespace mpe {

constexpr int kActorMaxD = 64;  // obs_dim bound (simple_spread N=12 has 28)
constexpr int kActorMaxA = 16;  // act0 + act1 bound (simple_reference has 5 + 10)
constexpr int kHid = 64;        // dense1 width / BiLSTM output width
constexpr int kH = 32;          // LSTM hidden size per direction
constexpr int kGateK = 96;      // [h1 (64) | h_prev (32)] contraction length of one LSTM step
constexpr int kGateN = 128;     // i, f, g, o gates x 32 units

// Device-resident packed weights.  Everything is stored k-major ("transposed") so that a thread's
// output columns are contiguous:  Wg[dir][k][col'] with col' = ug*8 + gate*2 + uu holding the
// reference row gate*32 + (ug + 16*uu) of [weight_ih | weight_hh]; biases b_ih + b_hh pre-summed.
// fp16 hi/lo weight image of the tensor-core path (byte offsets; see tc_kernels.cu)
struct TcDev {
  unsigned char *blob = nullptr;
  uint32_t bytes = 0;
  int D = 0, Kx = 0, A0 = 0, A1 = 0, A = 0;
  uint32_t off_wih[2][2] = {}, off_whh[2][2] = {}, off_w1[2] = {};
  uint32_t off_w2f = 0, off_bg = 0, off_b1 = 0, off_b2 = 0;
  float *scratch = nullptr;  // [148*2 pipelines][12 agents][16][128] forward-pass logits shares (teams of > 3 agents)
};

enum { kImplAuto = 0, kImplSimt = 1, kImplTc = 2, kImplTcFusedLarge = 3 };

struct ActorDev {
  TcDev tc;
  int impl = kImplAuto;
  float *blob = nullptr;
  size_t blob_floats = 0;
  size_t smem_floats = 0;  // leading part of the blob a CTA stages in shared memory (everything but the dense3 head)
  int D = 0, A0 = 0, A1 = 0, A = 0, Apad = 0, Dpad = 0, has_model = 0;
  int off_wg[2] = {0, 0};  // [96][128]
  int off_bg = 0;          // [2][128]
  int off_w1 = 0;          // [D][64]
  int off_b1 = 0;          // [64]
  int off_w2 = 0;          // [64][Apad]   (dense2 | dense2_1 ++ dense2_2)
  int off_b2 = 0;          // [Apad]
  int off_w3 = 0;          // [64][Dpad]   (dense3)
  int off_b3 = 0;          // [Dpad]
};

struct ActorHostWeights {
  const float *dense1_w, *dense1_b, *w_ih, *w_hh, *b_ih, *b_hh, *w_ih_r, *w_hh_r, *b_ih_r, *b_hh_r;
  const float *dense2_w, *dense2_b, *dense2b_w, *dense2b_b, *dense3_w, *dense3_b;
};

struct ActorIO {
  const float *obs = nullptr;     // [B][N][D]
  const float *gumbel = nullptr;  // [B][N][A] or null (Philox)
  float *logits = nullptr;        // [B][N][A]
  float *next_state = nullptr;    // [B][N][D]
  int32_t *act_u = nullptr, *act_c = nullptr;  // [B][N]
  float *onehot = nullptr;        // [B][N][A]
  float *hcat = nullptr;          // [B][N][64] relu(BiLSTM output): tensor-core path only, feeds launch_dense3
  int64_t B = 0, gid0 = 0;
  int32_t N = 0;
  uint64_t seed = 0, step = 0;
};

struct RolloutIO {
  int32_t T = 0;
  uint64_t step0 = 0;
  float *obs_next = nullptr;  // [T][B][N][D]
  float *rew = nullptr;       // [T][B][N]
  int32_t *act_u = nullptr, *act_c = nullptr;  // [T][B][N]
  float *obs_work = nullptr;  // [B][N][D] large teams: holds the current observations on entry, rewritten every step
};

void actor_layout(int D, int A0, int A1, bool has_model, ActorDev *out);
void actor_pack(const ActorDev &d, const ActorHostWeights &w, float *host_blob);
bool actor_supported(int N);
bool rollout_supported(int scenario, int N);
cudaError_t launch_actor_forward(const ActorDev &w, const ActorIO &io, cudaStream_t st);
// tensor-core path (tc_kernels.cu)
void tc_layout(int D, int A0, int A1, TcDev *out);
void tc_pack(const TcDev &t, const ActorHostWeights &w, unsigned char *host_image);
bool tc_actor_supported(const TcDev &w, int N);
bool tc_rollout_supported(const TcDev &w, int N);
size_t tc_scratch_floats(int sm_count);
cudaError_t launch_actor_forward_tc(const TcDev &w, const ActorIO &io, cudaStream_t st);
cudaError_t launch_rollout_tc(const EnvStateAny &env, const TcDev &w, const RolloutIO &io, cudaStream_t st);
// next_state = dense3(hcat): the "+model" head (ac_network_model_multi_gumbel.py:49,65) on top of the tensor-core forward
cudaError_t launch_dense3(const ActorDev &w, const float *hcat, int64_t rows, float *next_state, cudaStream_t st);
cudaError_t launch_rollout(const EnvStateAny &env, const ActorDev &w, const RolloutIO &io, cudaStream_t st);

}  // namespace mpe
