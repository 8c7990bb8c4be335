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
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402

B = 65536
NAG = int(os.environ.get('TL_N', '3'))  # TL_N=6: actor-only timeline of the large-team path
if NAG == 3:
    env = m.make_env('simple_spread', num_envs=B, batched=True, seed=1)
    actor = m.FusedActor(random_state_dict(10, 5, 1), seed=1)
    env.reset()
    for _ in range(3):
        env.rollout(actor, 1, record=True)
else:
    env = m.make_env('simple_spread', n=NAG, num_envs=B, batched=True, seed=1)
    actor = m.FusedActor(random_state_dict(env.obs_dim, 5, 1), seed=1)
    obs = env.reset()
    for _ in range(3):
        actor.forward(obs)
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

# whole-kernel view (k_tc2): globaltimer at entry / exit of every CTA, clock64 at entry, after the weight load,
# at the start of the CTA's second tile pair, at exit
g0, g1 = a[:, 2, 0], a[:, 2, 1]
ok = g0 > 0
print('kernel span over CTAs: %.1f us (first entry -> last exit); entry spread %.1f us' %
      ((g1[ok].max() - g0[ok].min()) / 1e3, (g0[ok].max() - g0[ok].min()) / 1e3))
c = a[:, 2, 2:6]
for cta in (0, 39, 40, 77, 107, 108, 147):
    e, wl, p2, x = c[cta]
    print('CTA %3d: weights ready +%d, second pair at %s, exit +%d cycles; wall %.1f us' %
          (cta, wl - e, ('+%d' % (p2 - e)) if p2 > 0 else '-', x - e, (g1[cta] - g0[cta]) / 1e3))

names = ['wait D1', 'ld d1', 'E1 math+st issue', 'wait st+arrive', 'wait G', 'ld va', 'chunk A', 'wait vb', 'chunk B+arrive', 'other']
for cta in (0, 77):
    for h in range(2):
        v = a[cta, 3, h * 10:h * 10 + 10]
        print('CTA %d WG%d phase cycles summed over the 13 half-iterations: ' % (cta, h) +
              ', '.join('%s %d' % (n, x) for n, x in zip(names, v)) + ' | total %d' % v.sum())
