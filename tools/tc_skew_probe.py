"""Experiment: de-phase the two warpgroups of k_tc2 (MPE_TC_SKEW = cycles warpgroup 1 waits before the cell
pipeline) so that the two warps of an SM sub-partition do not hit their TMEM round trips at the same time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402

B = 65536
env = m.make_env('simple_spread', num_envs=B, batched=True, seed=1)
actor = m.FusedActor(random_state_dict(10, 5, 1), seed=1)
env.reset()
for skew in [int(x) for x in (sys.argv[1:] or ['0', '300', '600', '1000', '1500', '2000', '3000', '0'])]:
    os.environ['MPE_TC_SKEW'] = str(skew)
    for _ in range(20):
        env.rollout(actor, 1, record=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        env.rollout(actor, 1, record=True)
    e1.record()
    torch.cuda.synchronize()
    print('skew %5d cycles: %.2f us per 65,536-env step (200 back-to-back launches)' % (skew, e0.elapsed_time(e1) * 5))
