"""Where the end-to-end (host-buffer) step goes: act_host and step_host timed separately, plus raw pinned copies."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402

B, N, D = 65536, 3, 10
env = m.make_env('simple_spread', num_envs=B, batched=True, seed=1, max_episode_len=25)
actor = m.FusedActor(random_state_dict(D, 5, 1), seed=1)
h_obs = torch.empty((B, N, D), dtype=torch.float32).pin_memory()
h_rew = torch.empty((B, N), dtype=torch.float32).pin_memory()
h_done = torch.empty((B, N), dtype=torch.uint8).pin_memory()
h_act = torch.empty((B, N), dtype=torch.int32).pin_memory()
h_obs.copy_(env.reset())
d_obs = torch.empty((B, N, D), device='cuda')


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print('act_host            %.1f us' % t(lambda: actor.act_host(h_obs, step=1, act_u=h_act)))
print('step_host           %.1f us' % t(lambda: env.step_host(h_act, out=(h_obs, h_rew, h_done))))
print('H2D obs 7.86 MB     %.1f us' % t(lambda: d_obs.copy_(h_obs, non_blocking=True)))
print('D2H obs 7.86 MB     %.1f us' % t(lambda: h_obs.copy_(d_obs, non_blocking=True)))
print('actor.forward (dev) %.1f us' % t(lambda: actor.forward(d_obs)))
