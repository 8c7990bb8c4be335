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
                      ('simple_speaker_listener', None, 129, 5)]:
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
    for impl in ['simt', 'tc']:
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
torch.cuda.synchronize()
print('sanitize target done')
