"""Small workload touching every kernel (ragged sizes) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402

for scen, n, B, A in [('simple_spread', None, 333, 5), ('simple_spread', 6, 70, 5), ('simple_spread', 9, 41, 5),
                      ('simple_spread', 12, 37, 5), ('simple_reference', None, 200, [5, 10]),
                      ('simple_speaker_listener', None, 129, 5), ('simple_spread', 1, 33, 5), ('simple_spread', 5, 65, 5),
                      ('simple_spread', 7, 29, 5), ('simple_spread', 8, 31, 5), ('simple_spread', 10, 27, 5),
                      ('simple_spread', 11, 23, 5), ('fullobs_collect_treasure', None, 77, 5)]:
    for precision in ('fp32', 'fp64'):
        env = m.make_env(scen, n=n, num_envs=B, batched=True, seed=3, precision=precision, max_episode_len=3)
        env.track_returns(True)
        env.reset()
        for t in range(4):
            au = torch.randint(0, 5, (B, env.n), dtype=torch.int32, device='cuda')
            ac = torch.randint(0, 10, (B, env.n), dtype=torch.int32, device='cuda') if env.act_c else None
            env.step(au, ac, info=True)
        env.reset(mask=torch.ones(B, dtype=torch.uint8))
        env.observe(); env.get_state(); env.read_stats()
    env = m.make_env(scen, n=n, num_envs=B, batched=True, seed=3, max_episode_len=3)
    obs = env.reset()
    tc_ok = env.n in (2, 3, 4, 6, 8, 9, 12)
    for impl in ['simt'] + (['tc'] if tc_ok else []) + (['tc_fused_large'] if env.n in (6, 9, 12) else []):
        actor = m.FusedActor(random_state_dict(env.obs_dim, A, 1), impl=impl)
        actor.forward(obs, want_logits=True, want_onehot=True)
        env.rollout(actor, 5, record=True)
        env.rollout(actor, 2)
# the tensor-core actor at tile-boundary batch sizes (1 tile, 1 tile + 1 row, odd numbers of tiles / tile pairs),
# small and large teams: same actions as the fp32 SIMT kernel under the same Philox keys except at near-ties
for n, D in ((3, 10), (4, 12), (6, 16), (12, 28)):
    sd = random_state_dict(D, 5, 2)
    tc, simt = m.FusedActor(sd, impl='tc', seed=4), m.FusedActor(sd, impl='simt', seed=4)
    for B in (1, 127, 128, 129, 255, 257, 300, 385):
        obs = torch.from_numpy(np.random.RandomState(B).uniform(-1, 1, (B, n, D)).astype(np.float32)).cuda()
        a, b = tc.forward(obs, step=B)['act_u'], simt.forward(obs, step=B)['act_u']
        assert int(a.min()) >= 0 and int(a.max()) <= 4
        assert float((a == b).float().mean()) > 0.995, (n, B)
# model head (dense3 on top of the tensor-core forward), critic, host-buffer rollouts at ragged sizes
for n, D, B in ((3, 10, 130), (6, 16, 257), (12, 28, 65)):
    sd = random_state_dict(D, 5, 3, model_head=True)
    obs = torch.from_numpy(np.random.RandomState(n).uniform(-1, 1, (B, n, D)).astype(np.float32)).cuda()
    for impl in ('tc', 'simt'):
        out = m.FusedActor(sd, impl=impl).forward(obs, want_next_state=True)
        assert bool(torch.isfinite(out['next_state']).all())
    rng = np.random.RandomState(0)
    csd = {'dense1.module.weight': rng.randn(64, D + 5).astype(np.float32) * 0.1, 'dense1.module.bias': np.zeros(64, np.float32),
           'lstm.weight_ih_l0': rng.randn(256, 64).astype(np.float32) * 0.1, 'lstm.weight_hh_l0': rng.randn(256, 64).astype(np.float32) * 0.1,
           'lstm.bias_ih_l0': np.zeros(256, np.float32), 'lstm.bias_hh_l0': np.zeros(256, np.float32),
           'dense2.weight': rng.randn(1, 64).astype(np.float32), 'dense2.bias': np.zeros(1, np.float32)}
    q = m.FusedCritic(csd, obs_dim=D).forward(obs, torch.eye(5, device='cuda')[torch.randint(0, 5, (B, n), device='cuda')])
    assert q.shape == (B, 1) and bool(torch.isfinite(q).all())
for scen, B, shards in (('simple_spread', 1001, 3), ('simple_reference', 130, 2), ('simple_spread', 5, 4)):
    A = [5, 10] if scen == 'simple_reference' else 5
    D = 21 if scen == 'simple_reference' else 10
    hr = m.HostRollout(scen, B, m.FusedActor(random_state_dict(D, A, 1), seed=2), shards=shards, seed=2, max_episode_len=3)
    for _ in range(7):
        hr.step(lambda k, tr: float(tr.rew_np.sum()))
    hr.flush()
    hr.close()
torch.cuda.synchronize()
print('sanitize target done')
