"""Back-to-back timing of the fused step kernel (mpe_rollout, T=1, 65,536 envs, simple_spread N=3): no L2 flush, 200
launches inside one CUDA-event pair, so launch latency and event overhead are amortised.  This is the number the
kernel experiments in profiles/README.md are quoted in; bench.py's per-step event timing with an L2 flush in between
comes out ~6 us higher.  A/B a variant build with MPE_B200_LIB=tmp_ab/libmpe_b200_<name>.so."""
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
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for _ in range(20):
        env.rollout(actor, 1, record=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        env.rollout(actor, 1, record=True)
    e1.record()
    torch.cuda.synchronize()
    print('%.2f us per 65,536-env step (200 back-to-back launches)' % (e0.elapsed_time(e1) * 5))
