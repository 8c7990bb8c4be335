"""Phase timeline of the tensor-core rollout kernel (first tile of two CTAs): python tools/tc_timeline.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ['MPE_TC_TIMELINE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200 import _lib  # noqa: E402
from oracle import actor_ref  # noqa: E402

B = 65536
env = m.make_env('simple_spread', num_envs=B, batched=True, seed=1)
actor = m.FusedActor(actor_ref.init_state_dict(10, 5, 1), seed=1)
env.reset()
for _ in range(3):
    env.rollout(actor, 1, record=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * (148 * 128))()
lib.mpe_debug_tc_timeline(buf, 148 * 128)
a = np.array(list(buf), dtype=np.int64).reshape(148, 4, 32)
names = {0: 'WG0', 1: 'WG1', 2: 'MMA0', 3: 'MMA1'}
for cta in (0, 77):
    t0 = a[cta, 0, 0]
    print('--- CTA', cta)
    for role in range(4):
        v = a[cta, role]
        print(names[role], ' '.join('%d:%d' % (i, v[i] - t0) for i in range(32) if v[i] > 0))
