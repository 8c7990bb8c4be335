// extern "C" surface of libmpe_b200.so (declared in include/mpe_b200.h).  Handles, argument
// checking, error text; all compute is in the kernel translation units behind env_launch.h /
// actor_launch.h.  No host synchronisation except where the header says so.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/mpe_b200.h"
#include "actor_launch.h"
#include "critic_launch.h"
#include "env_launch.h"
#include "replay_launch.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char *where) {
  g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return MPE_ECUDA;
}

#define CK(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return fail_cuda(e_, #call);  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true, switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev) {
      if (cudaSetDevice(dev) != cudaSuccess) ok = false; else switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);  // the common case - already on the handle's device - costs one cudaGetDevice
  }
};

size_t real_size(int precision) { return precision == MPE_F64 ? sizeof(double) : sizeof(float); }

}  // namespace

struct MpeEnv {
  mpe::EnvStateAny st;
  int device = 0;
  // lazily allocated device mirrors for the *_host entry points
  int32_t *h_act_u = nullptr, *h_act_c = nullptr;
  void *h_obs = nullptr, *h_rew = nullptr;
  uint8_t *h_done = nullptr;
  // scratch of the multi-kernel rollout (teams of > 3 agents): current observations and sampled actions
  float *r_obs = nullptr;
  int32_t *r_act = nullptr;
  // mpe_act_step_host_async: device mirror of the caller's observations and of its transition block
  float *hb_obs_in = nullptr;
  unsigned char *hb_block = nullptr;
  // dense2-share scratch of the tensor-core actor for launches made on THIS env's stream (mpe_rollout's per-step path,
  // mpe_act_step_host_async): several env shards share one actor handle on different streams (HostRollout), and the
  // large-team / two-head actor kernels park partial logits in a scratch buffer - one per stream, not one per actor
  float *tc_scratch = nullptr;
  bool synced = false;      // all envs are at the same episode step (true after an unmasked reset)
  int32_t host_tstep = 0;   // that common episode step
};

constexpr int kHostSlots = 4;  // independent sets of device mirrors: that many *_host_async calls may be in flight

struct ActorMirror {
  float *obs = nullptr, *onehot = nullptr;  // device mirrors for actor_forward_host[_async]
  int32_t *act_u = nullptr, *act_c = nullptr;
  float *tc_scratch = nullptr;  // slots > 0: own dense2-share scratch (slot 0 uses the actor's), see MpeEnv::tc_scratch
  int64_t cap = 0;  // capacity in rows (B*N)
};

struct MpeActor {
  mpe::ActorDev dev;
  int device = 0;
  ActorMirror mirror[kHostSlots];
  float *hcat = nullptr;  // [rows][64] relu(BiLSTM output) between the tensor-core forward and the dense3 head
  int64_t hcat_rows = 0;
};

struct MpeCritic {
  mpe::CriticDev dev;
  int device = 0;
};

struct MpeReplay {
  mpe::ReplayDev dev;
  int device = 0;
  int64_t size = 0, head = 0;  // host-side ring counters (rls/replay_buffer.py: len(_storage), _next_idx)
  uint64_t draws = 0;          // Philox counter of make_index calls
  int64_t *idx_scratch = nullptr;  // indices drawn by replay_sample when the caller passes none
  int64_t idx_cap = 0;
};

// The tensor-core actor of large teams (N > 3) and of two-head actors writes per-cell logits shares to a scratch
// buffer indexed by CTA: two launches in flight on different streams must not share it.
static bool tc_uses_scratch(const MpeActor *a, int N) { return N > 3 || a->dev.tc.A > 8; }
static cudaError_t ensure_tc_scratch(float **p, int device) {
  if (*p != nullptr) return cudaSuccess;
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  return cudaMalloc(p, mpe::tc_scratch_floats(nsm > 0 ? nsm : 148) * sizeof(float));
}

extern "C" {

int mpe_abi_version(void) { return MPE_ABI_VERSION; }
const char *mpe_last_error(void) { return g_err.c_str(); }

int mpe_create(const MpeConfig *cfg, MpeEnv **out) {
  if (cfg == nullptr || out == nullptr) return fail(MPE_EINVAL, "mpe_create: null argument");
  *out = nullptr;
  if (cfg->num_envs <= 0) return fail(MPE_EINVAL, "mpe_create: num_envs must be > 0");
  if (cfg->precision != MPE_F32 && cfg->precision != MPE_F64) return fail(MPE_EINVAL, "mpe_create: bad precision");
  int N = cfg->num_agents, L, D, dimc, act_c = 0;
  switch (cfg->scenario) {
    case MPE_SIMPLE_SPREAD:
      if (N == 0) N = 3;
      L = N; D = 4 + 2 * L; dimc = 2;
      break;
    case MPE_SIMPLE_REFERENCE:
      if (N != 0 && N != 2) return fail(MPE_EUNSUPPORTED, "simple_reference has exactly 2 agents");
      N = 2; L = 3; D = 21; dimc = 10; act_c = 10;
      break;
    case MPE_SIMPLE_SPEAKER_LISTENER:
      if (N != 0 && N != 2) return fail(MPE_EUNSUPPORTED, "simple_speaker_listener has exactly 2 agents");
      N = 2; L = 3; D = 11; dimc = 3;
      break;
    case MPE_COLLECT_TREASURE:  // MAAC fork: make_world() takes no agent count (8 agents: 6 collectors + 2 deposits)
      if (N != 0 && N != 8) return fail(MPE_EUNSUPPORTED, "fullobs_collect_treasure has exactly 8 agents");
      N = 8; L = 6; D = 30; dimc = 2;
      break;
    default:
      return fail(MPE_EUNSUPPORTED, "mpe_create: unknown scenario");
  }
  if (!mpe::env_supported(cfg->scenario, N)) {
    char buf[128];
    snprintf(buf, sizeof buf, "mpe_create: no kernel instantiated for scenario %d with %d agents", cfg->scenario, N);
    return fail(MPE_EUNSUPPORTED, buf);
  }
  DeviceGuard g(cfg->device);
  if (!g.ok) return fail(MPE_ECUDA, "mpe_create: cannot select device");
  MpeEnv *env = new (std::nothrow) MpeEnv();
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_create: out of host memory");
  env->device = cfg->device;
  mpe::EnvStateAny &s = env->st;
  s.B = cfg->num_envs; s.gid0 = cfg->env_id_offset; s.seed = cfg->seed;
  s.max_speed = cfg->max_speed; s.accel = cfg->accel;
  if (cfg->scenario == MPE_COLLECT_TREASURE) {  // the scenario sets accel = 1.5 and max_speed = 1.0 on every agent
    if (s.max_speed < 0.0) s.max_speed = 1.0;
    if (s.accel < 0.0) s.accel = 1.5;
  }
  s.precision = cfg->precision; s.scenario = cfg->scenario;
  s.N = N; s.L = L; s.D = D; s.dimc = dimc; s.act_u = 5; s.act_c = act_c;
  s.max_episode_len = cfg->max_episode_len > 0 ? cfg->max_episode_len : 0;
  const size_t rs = real_size(cfg->precision);
  const size_t B = (size_t)s.B;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void **p, size_t bytes, int fill) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, fill, bytes);
  };
  alloc(&s.pv, (size_t)N * B * 4 * rs, 0);
  alloc(&s.lm, (size_t)L * B * 2 * rs, 0);
  alloc(&s.ep_ret, B * rs, 0);
  alloc(reinterpret_cast<void **>(&s.goal), B * sizeof(int32_t), 0xFF);
  alloc(reinterpret_cast<void **>(&s.episode), B * sizeof(uint32_t), 0xFF);  // first reset -> episode 0
  alloc(reinterpret_cast<void **>(&s.tstep), B * sizeof(int32_t), 0);
  alloc(reinterpret_cast<void **>(&s.stats), 8 * sizeof(double), 0);  // MPE_STATS_LEN used, padded to 64 B
  if (cfg->scenario == MPE_SIMPLE_REFERENCE) alloc(&s.comm, (size_t)N * dimc * B * rs, 0);
  if (e != cudaSuccess) {
    mpe_destroy(env);
    return fail_cuda(e, "mpe_create: cudaMalloc");
  }
  *out = env;
  return MPE_OK;
}

int mpe_destroy(MpeEnv *env) {
  if (env == nullptr) return MPE_OK;
  DeviceGuard g(env->device);
  cudaDeviceSynchronize();
  mpe::EnvStateAny &s = env->st;
  void *ptrs[] = {s.pv, s.lm, s.ep_ret, s.comm, s.goal, s.episode, s.tstep, s.stats,
                  env->h_act_u, env->h_act_c, env->h_obs, env->h_rew, env->h_done, env->r_obs, env->r_act,
                  env->hb_obs_in, env->hb_block, env->tc_scratch};
  for (void *p : ptrs)
    if (p != nullptr) cudaFree(p);
  delete env;
  return MPE_OK;
}

int mpe_query(const MpeEnv *env, MpeDims *out) {
  if (env == nullptr || out == nullptr) return fail(MPE_EINVAL, "mpe_query: null argument");
  const mpe::EnvStateAny &s = env->st;
  out->num_agents = s.N; out->num_landmarks = s.L; out->obs_dim = s.D; out->dim_c = s.dimc;
  out->act_u = s.act_u; out->act_c = s.act_c; out->precision = s.precision; out->device = env->device;
  out->num_envs = s.B; out->env_id_offset = s.gid0;
  return MPE_OK;
}

int mpe_seed(MpeEnv *env, uint64_t seed, void *stream) {
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_seed: null env");
  DeviceGuard g(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mpe::EnvStateAny &s = env->st;
  s.seed = seed;
  // restart the episode counters so that (seed, env id, episode) streams are reproducible, and drop the
  // half-finished episodes' bookkeeping (the next reset starts episode 0 of the new seed)
  CK(cudaMemsetAsync(s.episode, 0xFF, (size_t)s.B * sizeof(uint32_t), st));
  CK(cudaMemsetAsync(s.tstep, 0, (size_t)s.B * sizeof(int32_t), st));
  CK(cudaMemsetAsync(s.ep_ret, 0, (size_t)s.B * real_size(s.precision), st));
  env->synced = false;
  env->host_tstep = 0;
  return MPE_OK;
}

int mpe_reset(MpeEnv *env, const uint8_t *mask, void *obs_out, void *stream) {
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_reset: null env");
  DeviceGuard g(env->device);
  CK(mpe::launch_reset(env->st, mask, obs_out, 0, static_cast<cudaStream_t>(stream)));
  env->synced = mask == nullptr;  // every env starts an episode together
  env->host_tstep = mask == nullptr ? 0 : env->host_tstep;
  return MPE_OK;
}

int mpe_set_state(MpeEnv *env, const void *pos, const void *vel, const void *lm, const int32_t *goal, void *stream) {
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_set_state: null env");
  DeviceGuard g(env->device);
  CK(mpe::launch_set_state(env->st, pos, vel, lm, goal, static_cast<cudaStream_t>(stream)));
  env->synced = false;
  return MPE_OK;
}

int mpe_get_state(MpeEnv *env, void *pos, void *vel, void *lm, int32_t *goal, void *stream) {
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_get_state: null env");
  DeviceGuard g(env->device);
  CK(mpe::launch_get_state(env->st, pos, vel, lm, goal, static_cast<cudaStream_t>(stream)));
  return MPE_OK;
}

int mpe_observe(MpeEnv *env, void *obs_out, void *stream) {
  if (env == nullptr || obs_out == nullptr) return fail(MPE_EINVAL, "mpe_observe: null argument");
  DeviceGuard g(env->device);
  CK(mpe::launch_observe(env->st, obs_out, static_cast<cudaStream_t>(stream)));
  return MPE_OK;
}

int mpe_step(MpeEnv *env, const int32_t *act_u, const int32_t *act_c, const void *comm_vec, void *obs, void *rew,
             uint8_t *done, int32_t *info_i, void *info_f, void *stream) {
  if (env == nullptr || act_u == nullptr) return fail(MPE_EINVAL, "mpe_step: null env or act_u");
  if (env->st.scenario == MPE_SIMPLE_REFERENCE && act_c == nullptr && comm_vec == nullptr)
    return fail(MPE_EINVAL, "mpe_step: simple_reference needs act_c or comm_vec");
  DeviceGuard g(env->device);
  CK(mpe::launch_step(env->st, act_u, act_c, comm_vec, obs, rew, done, info_i, info_f,
                      static_cast<cudaStream_t>(stream)));
  if (env->st.track) env->host_tstep += 1; else env->synced = false;
  return MPE_OK;
}

static int mpe_step_host_impl(MpeEnv *env, const int32_t *act_u_host, const int32_t *act_c_host, void *obs_host,
                              void *rew_host, uint8_t *done_host, cudaStream_t st, bool sync, const char *who) {
  if (env == nullptr || act_u_host == nullptr) return fail(MPE_EINVAL, std::string(who) + ": null env or act_u");
  if (env->st.scenario == MPE_SIMPLE_REFERENCE && act_c_host == nullptr)
    return fail(MPE_EINVAL, std::string(who) + ": simple_reference needs act_c");
  DeviceGuard g(env->device);
  const mpe::EnvStateAny &s = env->st;
  const size_t rows = (size_t)s.B * s.N, rs = real_size(s.precision);
  if (env->h_act_u == nullptr) {
    CK(cudaMalloc(&env->h_act_u, rows * sizeof(int32_t)));
    CK(cudaMalloc(&env->h_act_c, rows * sizeof(int32_t)));
    CK(cudaMalloc(&env->h_obs, rows * s.D * rs));
    CK(cudaMalloc(&env->h_rew, rows * rs));
    CK(cudaMalloc(&env->h_done, rows));
  }
  CK(cudaMemcpyAsync(env->h_act_u, act_u_host, rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (act_c_host != nullptr)
    CK(cudaMemcpyAsync(env->h_act_c, act_c_host, rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(mpe::launch_step(s, env->h_act_u, act_c_host != nullptr ? env->h_act_c : nullptr, nullptr,
                      obs_host != nullptr ? env->h_obs : nullptr, rew_host != nullptr ? env->h_rew : nullptr,
                      done_host != nullptr ? env->h_done : nullptr, nullptr, nullptr, st));
  if (env->st.track) env->host_tstep += 1; else env->synced = false;
  if (obs_host != nullptr) CK(cudaMemcpyAsync(obs_host, env->h_obs, rows * s.D * rs, cudaMemcpyDeviceToHost, st));
  if (rew_host != nullptr) CK(cudaMemcpyAsync(rew_host, env->h_rew, rows * rs, cudaMemcpyDeviceToHost, st));
  if (done_host != nullptr) CK(cudaMemcpyAsync(done_host, env->h_done, rows, cudaMemcpyDeviceToHost, st));
  if (sync) CK(cudaStreamSynchronize(st));
  return MPE_OK;
}

int mpe_step_host(MpeEnv *env, const int32_t *act_u_host, const int32_t *act_c_host, void *obs_host, void *rew_host,
                  uint8_t *done_host, void *stream) {
  return mpe_step_host_impl(env, act_u_host, act_c_host, obs_host, rew_host, done_host, static_cast<cudaStream_t>(stream),
                            true, "mpe_step_host");
}

int mpe_step_host_async(MpeEnv *env, const int32_t *act_u_host, const int32_t *act_c_host, void *obs_host,
                        void *rew_host, uint8_t *done_host, void *stream) {
  return mpe_step_host_impl(env, act_u_host, act_c_host, obs_host, rew_host, done_host, static_cast<cudaStream_t>(stream),
                            false, "mpe_step_host_async");
}

int mpe_reset_host_async(MpeEnv *env, void *obs_host, void *stream) {
  if (env == nullptr || obs_host == nullptr) return fail(MPE_EINVAL, "mpe_reset_host_async: null argument");
  DeviceGuard g(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mpe::EnvStateAny &s = env->st;
  const size_t rows = (size_t)s.B * s.N, rs = real_size(s.precision);
  if (env->h_obs == nullptr) {
    CK(cudaMalloc(&env->h_act_u, rows * sizeof(int32_t)));
    CK(cudaMalloc(&env->h_act_c, rows * sizeof(int32_t)));
    CK(cudaMalloc(&env->h_obs, rows * s.D * rs));
    CK(cudaMalloc(&env->h_rew, rows * rs));
    CK(cudaMalloc(&env->h_done, rows));
  }
  CK(mpe::launch_reset(s, nullptr, env->h_obs, 0, st));
  env->synced = true;
  env->host_tstep = 0;
  CK(cudaMemcpyAsync(obs_host, env->h_obs, rows * s.D * rs, cudaMemcpyDeviceToHost, st));
  return MPE_OK;
}

int mpe_track_returns(MpeEnv *env, int32_t enable) {
  if (env == nullptr) return fail(MPE_EINVAL, "mpe_track_returns: null env");
  env->st.track = enable ? 1 : 0;
  return MPE_OK;
}

int mpe_stats_read(MpeEnv *env, double out[MPE_STATS_LEN], int32_t clear, void *stream) {
  if (env == nullptr || out == nullptr) return fail(MPE_EINVAL, "mpe_stats_read: null argument");
  DeviceGuard g(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(out, env->st.stats, MPE_STATS_LEN * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (clear) CK(cudaMemsetAsync(env->st.stats, 0, MPE_STATS_LEN * sizeof(double), st));
  CK(cudaStreamSynchronize(st));
  return MPE_OK;
}

int mpe_stats_ptr(MpeEnv *env, double **dev_ptr) {
  if (env == nullptr || dev_ptr == nullptr) return fail(MPE_EINVAL, "mpe_stats_ptr: null argument");
  *dev_ptr = env->st.stats;
  return MPE_OK;
}

// ------------------------------------------------------------------------------------------------
// actor
// ------------------------------------------------------------------------------------------------
int actor_destroy(MpeActor *a);

int actor_create(const ActorConfig *cfg, MpeActor **out) {
  if (cfg == nullptr || out == nullptr) return fail(MPE_EINVAL, "actor_create: null argument");
  *out = nullptr;
  if (cfg->obs_dim <= 0 || cfg->obs_dim > mpe::kActorMaxD) return fail(MPE_EUNSUPPORTED, "actor_create: obs_dim out of range");
  if (cfg->act0 <= 0 || cfg->act1 < 0 || cfg->act0 + cfg->act1 > mpe::kActorMaxA)
    return fail(MPE_EUNSUPPORTED, "actor_create: head widths out of range");
  DeviceGuard g(cfg->device);
  if (!g.ok) return fail(MPE_ECUDA, "actor_create: cannot select device");
  MpeActor *a = new (std::nothrow) MpeActor();
  if (a == nullptr) return fail(MPE_EINVAL, "actor_create: out of host memory");
  a->device = cfg->device;
  mpe::actor_layout(cfg->obs_dim, cfg->act0, cfg->act1, cfg->has_model_head != 0, &a->dev);
  mpe::tc_layout(cfg->obs_dim, cfg->act0, cfg->act1, &a->dev.tc);
  cudaError_t e = cudaMalloc(&a->dev.blob, a->dev.blob_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(a->dev.blob, 0, a->dev.blob_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&a->dev.tc.blob, a->dev.tc.bytes);
  if (e == cudaSuccess) e = cudaMemset(a->dev.tc.blob, 0, a->dev.tc.bytes);
  if (e == cudaSuccess) {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device);
    e = cudaMalloc(&a->dev.tc.scratch, mpe::tc_scratch_floats(nsm > 0 ? nsm : 148) * sizeof(float));
  }
  if (e != cudaSuccess) {
    actor_destroy(a);  // frees whatever was allocated before the failure
    return fail_cuda(e, "actor_create: cudaMalloc");
  }
  *out = a;
  return MPE_OK;
}

int actor_destroy(MpeActor *a) {
  if (a == nullptr) return MPE_OK;
  DeviceGuard g(a->device);
  cudaDeviceSynchronize();
  void *ptrs[] = {a->dev.blob, a->dev.tc.blob, a->dev.tc.scratch, a->hcat};
  for (void *p : ptrs)
    if (p != nullptr) cudaFree(p);
  for (ActorMirror &m : a->mirror) {
    void *mp[] = {m.obs, m.onehot, m.act_u, m.act_c, m.tc_scratch};
    for (void *p : mp)
      if (p != nullptr) cudaFree(p);
  }
  delete a;
  return MPE_OK;
}

int actor_load(MpeActor *a, const ActorWeights *w, void *stream) {
  if (a == nullptr || w == nullptr) return fail(MPE_EINVAL, "actor_load: null argument");
  if (!w->dense1_w || !w->dense1_b || !w->w_ih || !w->w_hh || !w->b_ih || !w->b_hh || !w->w_ih_r || !w->w_hh_r ||
      !w->b_ih_r || !w->b_hh_r || !w->dense2_w || !w->dense2_b)
    return fail(MPE_EINVAL, "actor_load: missing tensor");
  if (a->dev.A1 > 0 && (!w->dense2b_w || !w->dense2b_b)) return fail(MPE_EINVAL, "actor_load: missing dense2_2");
  if (a->dev.has_model && (!w->dense3_w || !w->dense3_b)) return fail(MPE_EINVAL, "actor_load: missing dense3");
  DeviceGuard g(a->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float *host = nullptr;
  CK(cudaMallocHost(&host, a->dev.blob_floats * sizeof(float)));
  mpe::ActorHostWeights hw = {w->dense1_w, w->dense1_b, w->w_ih,     w->w_hh,     w->b_ih,      w->b_hh,
                              w->w_ih_r,   w->w_hh_r,   w->b_ih_r,   w->b_hh_r,   w->dense2_w,  w->dense2_b,
                              w->dense2b_w, w->dense2b_b, w->dense3_w, w->dense3_b};
  mpe::actor_pack(a->dev, hw, host);
  cudaError_t e = cudaMemcpyAsync(a->dev.blob, host, a->dev.blob_floats * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the pinned staging buffer is freed below
  cudaFreeHost(host);
  if (e == cudaSuccess) {  // fp16 hi/lo image of the tensor-core path
    unsigned char *img = nullptr;
    e = cudaMallocHost(&img, a->dev.tc.bytes);
    if (e == cudaSuccess) {
      mpe::tc_pack(a->dev.tc, hw, img);
      e = cudaMemcpyAsync(a->dev.tc.blob, img, a->dev.tc.bytes, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      cudaFreeHost(img);
    }
  }
  if (e != cudaSuccess) return fail_cuda(e, "actor_load: upload");
  return MPE_OK;
}

static bool use_tc(const MpeActor *a, int N, bool needs_simt) {
  if (a->dev.impl == mpe::kImplSimt || needs_simt) return false;
  return mpe::tc_actor_supported(a->dev.tc, N);
}

int actor_set_impl(MpeActor *a, int32_t impl) {
  if (a == nullptr || impl < 0 || impl > 3) return fail(MPE_EINVAL, "actor_set_impl: bad argument");
  a->dev.impl = impl;
  return MPE_OK;
}

int actor_forward(MpeActor *a, const float *obs, int64_t B, int32_t N, const float *gumbel, uint64_t seed,
                  uint64_t step, int64_t env_id_offset, float *logits, float *next_state, int32_t *act_u,
                  int32_t *act_c, float *onehot, void *stream) {
  if (a == nullptr || obs == nullptr) return fail(MPE_EINVAL, "actor_forward: null argument");
  if (B <= 0) return MPE_OK;
  if (!mpe::actor_supported(N)) return fail(MPE_EUNSUPPORTED, "actor_forward: unsupported agent count");
  if (next_state != nullptr && !a->dev.has_model) return fail(MPE_EINVAL, "actor_forward: no model head loaded");
  DeviceGuard g(a->device);
  mpe::ActorIO io;
  io.obs = obs; io.gumbel = gumbel; io.logits = logits; io.next_state = next_state;
  io.act_u = act_u; io.act_c = act_c; io.onehot = onehot;
  io.B = B; io.N = N; io.seed = seed; io.step = step; io.gid0 = env_id_offset;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((a->dev.impl == mpe::kImplTc || a->dev.impl == mpe::kImplTcFusedLarge) && !use_tc(a, N, false))
    return fail(MPE_EUNSUPPORTED, "actor_forward: the tensor-core path covers 2/3/4/6/9/12 agents, obs_dim <= 32, <= 8 head entries for > 3 agents");
  if (use_tc(a, N, false)) {
    if (next_state != nullptr) {  // "+model" head: the forward leaves relu(hcat) behind, one small kernel applies dense3
      const int64_t rows = B * N;
      if (rows > a->hcat_rows) {
        if (a->hcat != nullptr) { CK(cudaStreamSynchronize(st)); cudaFree(a->hcat); a->hcat = nullptr; a->hcat_rows = 0; }
        CK(cudaMalloc(&a->hcat, (size_t)rows * 64 * sizeof(float)));
        a->hcat_rows = rows;
      }
      io.hcat = a->hcat;
      io.next_state = nullptr;
    }
    CK(mpe::launch_actor_forward_tc(a->dev.tc, io, st));
    if (next_state != nullptr) CK(mpe::launch_dense3(a->dev, a->hcat, B * N, next_state, st));
  } else {
    CK(mpe::launch_actor_forward(a->dev, io, st));
  }
  return MPE_OK;
}

// H2D obs -> forward + sample -> D2H actions, all enqueued on `st`; `sync` waits for the stream at the end.
static int actor_forward_host_impl(MpeActor *a, const float *obs_host, int64_t B, int32_t N, uint64_t seed, uint64_t step,
                                   int64_t env_id_offset, int32_t *act_u_host, int32_t *act_c_host, float *onehot_host,
                                   int32_t slot, cudaStream_t st, bool sync, const char *who) {
  if (a == nullptr || obs_host == nullptr) return fail(MPE_EINVAL, std::string(who) + ": null argument");
  if (slot < 0 || slot >= kHostSlots) return fail(MPE_EINVAL, std::string(who) + ": slot out of range");
  if (B <= 0) return MPE_OK;
  if (!mpe::actor_supported(N)) return fail(MPE_EUNSUPPORTED, std::string(who) + ": unsupported agent count");
  DeviceGuard g(a->device);
  const int64_t rows = B * N;
  const int A = a->dev.A0 + a->dev.A1;
  ActorMirror &m = a->mirror[slot];
  if (rows > m.cap) {
    // growing a mirror frees the old one: wait for whatever the stream still has in flight on it
    if (m.cap > 0) CK(cudaStreamSynchronize(st));
    void *old[] = {m.obs, m.onehot, m.act_u, m.act_c};
    for (void *p : old)
      if (p != nullptr) cudaFree(p);
    float *keep = m.tc_scratch;
    m = ActorMirror();
    m.tc_scratch = keep;
    CK(cudaMalloc(&m.obs, (size_t)rows * a->dev.D * sizeof(float)));
    CK(cudaMalloc(&m.onehot, (size_t)rows * A * sizeof(float)));
    CK(cudaMalloc(&m.act_u, (size_t)rows * sizeof(int32_t)));
    CK(cudaMalloc(&m.act_c, (size_t)rows * sizeof(int32_t)));
    m.cap = rows;
  }
  CK(cudaMemcpyAsync(m.obs, obs_host, (size_t)rows * a->dev.D * sizeof(float), cudaMemcpyHostToDevice, st));
  mpe::ActorIO io;
  io.obs = m.obs; io.act_u = m.act_u; io.act_c = a->dev.A1 > 0 ? m.act_c : nullptr;
  io.onehot = onehot_host != nullptr ? m.onehot : nullptr;
  io.B = B; io.N = N; io.seed = seed; io.step = step; io.gid0 = env_id_offset;
  if (use_tc(a, N, false)) {
    mpe::TcDev tc = a->dev.tc;
    if (slot > 0 && tc_uses_scratch(a, N)) {  // slots are in flight on different streams: own scratch each
      CK(ensure_tc_scratch(&m.tc_scratch, a->device));
      tc.scratch = m.tc_scratch;
    }
    CK(mpe::launch_actor_forward_tc(tc, io, st));
  } else {
    CK(mpe::launch_actor_forward(a->dev, io, st));
  }
  if (act_u_host != nullptr) CK(cudaMemcpyAsync(act_u_host, m.act_u, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (act_c_host != nullptr && a->dev.A1 > 0)
    CK(cudaMemcpyAsync(act_c_host, m.act_c, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (onehot_host != nullptr)
    CK(cudaMemcpyAsync(onehot_host, m.onehot, (size_t)rows * A * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (sync) CK(cudaStreamSynchronize(st));
  return MPE_OK;
}

int actor_forward_host(MpeActor *a, const float *obs_host, int64_t B, int32_t N, uint64_t seed, uint64_t step,
                       int64_t env_id_offset, int32_t *act_u_host, int32_t *act_c_host, float *onehot_host,
                       void *stream) {
  return actor_forward_host_impl(a, obs_host, B, N, seed, step, env_id_offset, act_u_host, act_c_host, onehot_host, 0,
                                 static_cast<cudaStream_t>(stream), true, "actor_forward_host");
}

int actor_forward_host_async(MpeActor *a, const float *obs_host, int64_t B, int32_t N, uint64_t seed, uint64_t step,
                             int64_t env_id_offset, int32_t *act_u_host, int32_t *act_c_host, float *onehot_host,
                             int32_t slot, void *stream) {
  return actor_forward_host_impl(a, obs_host, B, N, seed, step, env_id_offset, act_u_host, act_c_host, onehot_host, slot,
                                 static_cast<cudaStream_t>(stream), false, "actor_forward_host_async");
}

static MpeHostBlockLayout block_layout(const mpe::EnvStateAny &s) {
  const uint64_t rows = (uint64_t)s.B * s.N, rs = real_size(s.precision);
  auto up = [](uint64_t x) { return (x + 255) / 256 * 256; };
  MpeHostBlockLayout l;
  l.off_act_u = 0;
  l.off_act_c = up(l.off_act_u + rows * 4);
  l.off_obs = up(l.off_act_c + (s.act_c > 0 ? rows * 4 : 0));  // no message head: act_c has zero length (off_act_c == off_obs)
  l.off_rew = up(l.off_obs + rows * s.D * rs);
  l.off_done = up(l.off_rew + rows * rs);
  l.bytes = up(l.off_done + rows);
  return l;
}

int mpe_host_block_layout(const MpeEnv *env, MpeHostBlockLayout *out) {
  if (env == nullptr || out == nullptr) return fail(MPE_EINVAL, "mpe_host_block_layout: null argument");
  *out = block_layout(env->st);
  return MPE_OK;
}

int mpe_act_step_host_async(MpeEnv *env, MpeActor *actor, const float *obs_host, uint64_t step, void *block_host,
                            void *stream) {
  if (env == nullptr || actor == nullptr || obs_host == nullptr || block_host == nullptr)
    return fail(MPE_EINVAL, "mpe_act_step_host_async: null argument");
  mpe::EnvStateAny &s = env->st;
  if (s.precision != MPE_F32) return fail(MPE_EUNSUPPORTED, "mpe_act_step_host_async: fp32 envs only");
  if (env->device != actor->device) return fail(MPE_EINVAL, "mpe_act_step_host_async: env and actor on different devices");
  if (actor->dev.D != s.D || actor->dev.A0 != 5 || actor->dev.A1 != s.act_c)
    return fail(MPE_EINVAL, "mpe_act_step_host_async: actor does not match the env's observation / action space");
  if (!mpe::actor_supported(s.N)) return fail(MPE_EUNSUPPORTED, "mpe_act_step_host_async: unsupported agent count");
  DeviceGuard g(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MpeHostBlockLayout l = block_layout(s);
  const size_t rows = (size_t)s.B * s.N;
  if (env->hb_block == nullptr) {
    CK(cudaMalloc(&env->hb_obs_in, rows * s.D * sizeof(float)));
    CK(cudaMalloc(&env->hb_block, l.bytes));
    CK(cudaMemsetAsync(env->hb_block, 0, l.bytes, st));
  }
  unsigned char *blk = env->hb_block;
  CK(cudaMemcpyAsync(env->hb_obs_in, obs_host, rows * s.D * sizeof(float), cudaMemcpyHostToDevice, st));
  mpe::ActorIO io;
  io.obs = env->hb_obs_in;
  io.act_u = reinterpret_cast<int32_t *>(blk + l.off_act_u);
  io.act_c = actor->dev.A1 > 0 ? reinterpret_cast<int32_t *>(blk + l.off_act_c) : nullptr;
  io.B = s.B; io.N = s.N; io.seed = s.seed; io.step = step; io.gid0 = s.gid0;
  if (use_tc(actor, s.N, false)) {
    mpe::TcDev tc = actor->dev.tc;
    if (tc_uses_scratch(actor, s.N)) {  // shards share the actor handle on different streams: scratch per env handle
      CK(ensure_tc_scratch(&env->tc_scratch, env->device));
      tc.scratch = env->tc_scratch;
    }
    CK(mpe::launch_actor_forward_tc(tc, io, st));
  } else {
    CK(mpe::launch_actor_forward(actor->dev, io, st));
  }
  CK(mpe::launch_step(s, io.act_u, io.act_c, nullptr, blk + l.off_obs, blk + l.off_rew, blk + l.off_done, nullptr, nullptr, st));
  if (s.track) env->host_tstep += 1; else env->synced = false;
  CK(cudaMemcpyAsync(block_host, blk, l.bytes, cudaMemcpyDeviceToHost, st));
  return MPE_OK;
}

int mpe_host_alloc(void **out, uint64_t bytes) {
  if (out == nullptr || bytes == 0) return fail(MPE_EINVAL, "mpe_host_alloc: bad argument");
  *out = nullptr;
  CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
  return MPE_OK;
}

int mpe_host_free(void *p) {
  if (p != nullptr) CK(cudaFreeHost(p));
  return MPE_OK;
}

int mpe_host_wait(void *stream) {
  CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return MPE_OK;
}

int mpe_rollout(MpeEnv *env, MpeActor *actor, int32_t T, uint64_t step0, float *obs_next, float *rew, int32_t *act_u,
                int32_t *act_c, void *stream) {
  if (env == nullptr || actor == nullptr) return fail(MPE_EINVAL, "mpe_rollout: null handle");
  if (T <= 0) return MPE_OK;
  if (env->st.precision != MPE_F32) return fail(MPE_EUNSUPPORTED, "mpe_rollout: fp32 envs only");
  if (env->device != actor->device) return fail(MPE_EINVAL, "mpe_rollout: env and actor on different devices");
  if (actor->dev.D != env->st.D) return fail(MPE_EINVAL, "mpe_rollout: actor obs_dim does not match the env");
  if (actor->dev.A0 != 5 || actor->dev.A1 != env->st.act_c)
    return fail(MPE_EINVAL, "mpe_rollout: actor heads do not match the env's action space");
  if (!mpe::rollout_supported(env->st.scenario, env->st.N))
    return fail(MPE_EUNSUPPORTED, "mpe_rollout: unsupported scenario / agent count");
  DeviceGuard g(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mpe::RolloutIO io;
  io.T = T; io.step0 = step0; io.obs_next = obs_next; io.rew = rew; io.act_u = act_u; io.act_c = act_c;
  // One kernel for all T steps: teams of <= 3.  For the large simple_spread teams (6 / 9 / 12) the single-kernel form
  // exists too but is 17 - 20 % SLOWER than actor + step kernels per step (measured, DESIGN.md 4.3: the env step wants
  // the whole GPU's occupancy, inside the persistent actor kernel it runs on 8 warps per SM), so it is opt-in.
  if (actor->dev.impl != mpe::kImplSimt && mpe::tc_rollout_supported(actor->dev.tc, env->st.N) &&
      (env->st.N <= 3 || actor->dev.impl == mpe::kImplTcFusedLarge)) {
    if (env->st.N > 3) {  // large teams: the kernel streams the observations from a work buffer it rewrites every step
      const int64_t rows = env->st.B * env->st.N;
      if (env->r_obs == nullptr) {
        CK(cudaMalloc(&env->r_obs, (size_t)rows * env->st.D * sizeof(float)));
        CK(cudaMalloc(&env->r_act, (size_t)rows * 2 * sizeof(int32_t)));
      }
      CK(mpe::launch_observe(env->st, env->r_obs, st));
      io.obs_work = env->r_obs;
    }
    mpe::TcDev tcw = actor->dev.tc;
    if (tc_uses_scratch(actor, env->st.N)) {
      CK(ensure_tc_scratch(&env->tc_scratch, env->device));
      tcw.scratch = env->tc_scratch;
    }
    CK(mpe::launch_rollout_tc(env->st, tcw, io, st));
    env->synced = false;
    return MPE_OK;
  }
#ifdef MPE_AB_KERNELS
  if (actor->dev.impl == mpe::kImplSimt) {  // A/B builds: the fp32 FFMA fused kernel
    CK(mpe::launch_rollout(env->st, actor->dev, io, st));
    env->synced = false;
    return MPE_OK;
  }
#endif
  {
    // Every other combination (teams of > 3 agents, the fp32 FFMA actor): the same loop as kernels per step - actor
    // (tensor cores where the shape is covered, else FFMA), the env step kernel, and a reset kernel on the steps
    // where an episode can end.  Same Philox keys, same arithmetic: bit-identical to the stepwise calls.
    mpe::EnvStateAny s = env->st;
    s.track = 1;
    const int64_t rows = s.B * s.N;
    if (env->r_obs == nullptr) {
      CK(cudaMalloc(&env->r_obs, (size_t)rows * s.D * sizeof(float)));
      CK(cudaMalloc(&env->r_act, (size_t)rows * 2 * sizeof(int32_t)));
    }
    const bool tc = use_tc(actor, s.N, false);
    mpe::TcDev tcw = actor->dev.tc;
    if (tc && tc_uses_scratch(actor, s.N)) {  // rollouts of several envs may share one actor on different streams
      CK(ensure_tc_scratch(&env->tc_scratch, env->device));
      tcw.scratch = env->tc_scratch;
    }
    CK(mpe::launch_observe(s, env->r_obs, st));
    const float *cur = env->r_obs;
    const int L = s.max_episode_len;
    for (int t = 0; t < T; ++t) {
      mpe::ActorIO aio;
      aio.obs = cur; aio.B = s.B; aio.N = s.N; aio.seed = s.seed; aio.step = step0 + (uint64_t)t; aio.gid0 = s.gid0;
      aio.act_u = act_u != nullptr ? act_u + (int64_t)t * rows : env->r_act;
      if (s.act_c > 0) aio.act_c = act_c != nullptr ? act_c + (int64_t)t * rows : env->r_act + rows;
      if (tc)
        CK(mpe::launch_actor_forward_tc(tcw, aio, st));
      else
        CK(mpe::launch_actor_forward(actor->dev, aio, st));
      float *o = obs_next != nullptr ? obs_next + (int64_t)t * rows * s.D : env->r_obs;
      CK(mpe::launch_step(s, aio.act_u, aio.act_c, nullptr, o, rew != nullptr ? rew + (int64_t)t * rows : nullptr, nullptr,
                          nullptr, nullptr, st));
      cur = o;
      env->host_tstep += 1;
      const bool may_end = L > 0 && (!env->synced || env->host_tstep >= L);
      if (may_end) {  // experiments/run.py:59-60; also re-emits the observations the next actor call reads
        CK(mpe::launch_reset(s, nullptr, env->r_obs, L, st));
        cur = env->r_obs;
        if (env->synced) env->host_tstep = 0;
      }
    }
  }
  return MPE_OK;
}

// ------------------------------------------------------------------------------------------------
// critic
// ------------------------------------------------------------------------------------------------
int critic_destroy(MpeCritic *c);

int critic_create(const CriticConfig *cfg, MpeCritic **out) {
  if (cfg == nullptr || out == nullptr) return fail(MPE_EINVAL, "critic_create: null argument");
  *out = nullptr;
  if (cfg->obs_dim <= 0 || cfg->act_dim <= 0 || cfg->obs_dim + cfg->act_dim > mpe::kCriticMaxIn)
    return fail(MPE_EUNSUPPORTED, "critic_create: obs_dim + act_dim out of range");
  if (cfg->out_dim <= 0 || cfg->out_dim > mpe::kCriticMaxOut) return fail(MPE_EUNSUPPORTED, "critic_create: out_dim out of range");
  DeviceGuard g(cfg->device);
  if (!g.ok) return fail(MPE_ECUDA, "critic_create: cannot select device");
  MpeCritic *c = new (std::nothrow) MpeCritic();
  if (c == nullptr) return fail(MPE_EINVAL, "critic_create: out of host memory");
  c->device = cfg->device;
  mpe::critic_layout(cfg->obs_dim, cfg->act_dim, cfg->out_dim, cfg->has_reward_head != 0, cfg->relu_attention != 0, &c->dev);
  cudaError_t e = cudaMalloc(&c->dev.blob, c->dev.blob_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(c->dev.blob, 0, c->dev.blob_floats * sizeof(float));
  if (e != cudaSuccess) {
    critic_destroy(c);
    return fail_cuda(e, "critic_create: cudaMalloc");
  }
  *out = c;
  return MPE_OK;
}

int critic_destroy(MpeCritic *c) {
  if (c == nullptr) return MPE_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  if (c->dev.blob != nullptr) cudaFree(c->dev.blob);
  delete c;
  return MPE_OK;
}

int critic_load(MpeCritic *c, const CriticWeights *w, void *stream) {
  if (c == nullptr || w == nullptr) return fail(MPE_EINVAL, "critic_load: null argument");
  if (!w->dense1_w || !w->dense1_b || !w->w_ih || !w->w_hh || !w->b_ih || !w->b_hh || !w->dense2_w || !w->dense2_b)
    return fail(MPE_EINVAL, "critic_load: missing tensor");
  if (c->dev.has_r && (!w->dense3_w || !w->dense3_b)) return fail(MPE_EINVAL, "critic_load: missing dense3");
  DeviceGuard g(c->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float *host = nullptr;
  CK(cudaMallocHost(&host, c->dev.blob_floats * sizeof(float)));
  mpe::CriticHostWeights hw = {w->dense1_w, w->dense1_b, w->w_ih, w->w_hh, w->b_ih, w->b_hh,
                               w->dense2_w, w->dense2_b, w->dense3_w, w->dense3_b};
  mpe::critic_pack(c->dev, hw, host);
  cudaError_t e = cudaMemcpyAsync(c->dev.blob, host, c->dev.blob_floats * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the pinned staging buffer is freed below
  cudaFreeHost(host);
  if (e != cudaSuccess) return fail_cuda(e, "critic_load: upload");
  return MPE_OK;
}

int critic_forward(MpeCritic *c, const float *obs, const float *action, int64_t B, int32_t N, float *q, float *r,
                   void *stream) {
  if (c == nullptr || obs == nullptr || action == nullptr || q == nullptr) return fail(MPE_EINVAL, "critic_forward: null argument");
  if (B <= 0) return MPE_OK;
  if (N < 1 || N > mpe::kCriticMaxAgents) return fail(MPE_EUNSUPPORTED, "critic_forward: 1..16 agents");
  if (r != nullptr && !c->dev.has_r) return fail(MPE_EINVAL, "critic_forward: no reward head loaded");
  DeviceGuard g(c->device);
  CK(mpe::launch_critic_forward(c->dev, obs, action, B, N, q, r, static_cast<cudaStream_t>(stream)));
  return MPE_OK;
}

// ------------------------------------------------------------------------------------------------
// replay ring
// ------------------------------------------------------------------------------------------------
int replay_create(const ReplayConfig *cfg, MpeReplay **out) {
  if (cfg == nullptr || out == nullptr) return fail(MPE_EINVAL, "replay_create: null argument");
  *out = nullptr;
  if (cfg->capacity <= 0 || cfg->num_agents <= 0 || cfg->num_agents > 32 || cfg->obs_dim <= 0 || cfg->act0 <= 0 ||
      cfg->act0 > 127 || cfg->act1 < 0 || cfg->act1 > 127)
    return fail(MPE_EINVAL, "replay_create: bad configuration");
  DeviceGuard g(cfg->device);
  if (!g.ok) return fail(MPE_ECUDA, "replay_create: cannot select device");
  MpeReplay *r = new (std::nothrow) MpeReplay();
  if (r == nullptr) return fail(MPE_EINVAL, "replay_create: out of host memory");
  r->device = cfg->device;
  mpe::ReplayDev &d = r->dev;
  d.capacity = cfg->capacity; d.N = cfg->num_agents; d.D = cfg->obs_dim; d.A0 = cfg->act0; d.A1 = cfg->act1;
  d.rec_bytes = mpe::ReplayDev::record_bytes(d.N, d.D);
  cudaError_t e = cudaMalloc(&d.ring, (size_t)d.capacity * (size_t)d.rec_bytes);
  if (e != cudaSuccess) {
    replay_destroy(r);
    return fail_cuda(e, "replay_create: cudaMalloc");
  }
  *out = r;
  return MPE_OK;
}

int replay_destroy(MpeReplay *r) {
  if (r == nullptr) return MPE_OK;
  DeviceGuard g(r->device);
  cudaDeviceSynchronize();
  void *ptrs[] = {r->dev.ring, r->idx_scratch};
  for (void *p : ptrs)
    if (p != nullptr) cudaFree(p);
  delete r;
  return MPE_OK;
}

int replay_clear(MpeReplay *r) {
  if (r == nullptr) return fail(MPE_EINVAL, "replay_clear: null handle");
  r->size = 0; r->head = 0;
  return MPE_OK;
}

int64_t replay_len(const MpeReplay *r) { return r == nullptr ? 0 : r->size; }
int64_t replay_next_idx(const MpeReplay *r) { return r == nullptr ? 0 : r->head; }

int replay_add(MpeReplay *r, const float *obs, const int32_t *act_u, const int32_t *act_c, const float *rew,
               const float *obs_next, const float *done, int64_t B, void *stream) {
  if (r == nullptr || obs == nullptr || act_u == nullptr || rew == nullptr || obs_next == nullptr)
    return fail(MPE_EINVAL, "replay_add: null argument");
  if (B <= 0) return MPE_OK;
  if (B > r->dev.capacity) return fail(MPE_EINVAL, "replay_add: batch larger than the ring");
  if (r->dev.A1 > 0 && act_c == nullptr) return fail(MPE_EINVAL, "replay_add: act_c required for a two-head actor");
  DeviceGuard g(r->device);
  CK(mpe::launch_replay_add(r->dev, r->head, B, obs, act_u, act_c, rew, obs_next, done, static_cast<cudaStream_t>(stream)));
  r->head = (r->head + B) % r->dev.capacity;
  r->size = r->size + B < r->dev.capacity ? r->size + B : r->dev.capacity;
  return MPE_OK;
}

int replay_sample(MpeReplay *r, int64_t batch, const int64_t *idx, uint64_t seed, float *obs, float *act_onehot,
                  float *rew, float *obs_next, float *done, int64_t *idx_out, void *stream) {
  if (r == nullptr) return fail(MPE_EINVAL, "replay_sample: null handle");
  if (batch <= 0) return MPE_OK;
  if (r->size <= 0) return fail(MPE_EINVAL, "replay_sample: the ring is empty");
  DeviceGuard g(r->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t *use = idx;
  if (idx == nullptr) {  // ReplayBuffer.make_index: uniform with replacement, into idx_out or a private scratch
    int64_t *dst = idx_out;
    if (dst == nullptr) {
      if (batch > r->idx_cap) {
        if (r->idx_scratch != nullptr) cudaFree(r->idx_scratch);
        r->idx_scratch = nullptr; r->idx_cap = 0;
        CK(cudaMalloc(&r->idx_scratch, (size_t)batch * sizeof(int64_t)));
        r->idx_cap = batch;
      }
      dst = r->idx_scratch;
    }
    CK(mpe::launch_replay_make_index(r->size, batch, seed, r->draws, dst, st));
    use = dst;
  } else if (idx_out != nullptr) {
    CK(cudaMemcpyAsync(idx_out, idx, (size_t)batch * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  }
  if (obs != nullptr || act_onehot != nullptr || rew != nullptr || obs_next != nullptr || done != nullptr)
    CK(mpe::launch_replay_gather(r->dev, r->size, batch, use, obs, act_onehot, rew, obs_next, done, st));
  if (idx == nullptr) r->draws += 1;
  return MPE_OK;
}

}  // extern "C"

// precision dispatch for the env launchers declared in env_launch.h
namespace mpe {
bool env_supported(int scenario, int N) {
  if (scenario == 0) return N >= 1 && N <= 12;  // simple_spread: make_world(num_agents=n), experiments/scenarios.py:170
  if (scenario == 3) return N == 8;  // fullobs_collect_treasure
  return (scenario == 1 || scenario == 2) && N == 2;
}
cudaError_t launch_reset(const EnvStateAny &a, const uint8_t *mask, void *obs, int auto_len, cudaStream_t st) {
  return a.precision == MPE_F64 ? launch_reset_f64(a, mask, obs, auto_len, st) : launch_reset_f32(a, mask, obs, auto_len, st);
}
cudaError_t launch_observe(const EnvStateAny &a, void *obs, cudaStream_t st) {
  return a.precision == MPE_F64 ? launch_observe_f64(a, obs, st) : launch_observe_f32(a, obs, st);
}
cudaError_t launch_step(const EnvStateAny &a, const int32_t *act_u, const int32_t *act_c, const void *comm_vec,
                        void *obs, void *rew, uint8_t *done, int32_t *info_i, void *info_f, cudaStream_t st) {
  return a.precision == MPE_F64 ? launch_step_f64(a, act_u, act_c, comm_vec, obs, rew, done, info_i, info_f, st)
                                : launch_step_f32(a, act_u, act_c, comm_vec, obs, rew, done, info_i, info_f, st);
}
cudaError_t launch_set_state(const EnvStateAny &a, const void *pos, const void *vel, const void *lm,
                             const int32_t *goal, cudaStream_t st) {
  return a.precision == MPE_F64 ? launch_set_state_f64(a, pos, vel, lm, goal, st)
                                : launch_set_state_f32(a, pos, vel, lm, goal, st);
}
cudaError_t launch_get_state(const EnvStateAny &a, void *pos, void *vel, void *lm, int32_t *goal, cudaStream_t st) {
  return a.precision == MPE_F64 ? launch_get_state_f64(a, pos, vel, lm, goal, st)
                                : launch_get_state_f32(a, pos, vel, lm, goal, st);
}
}  // namespace mpe
