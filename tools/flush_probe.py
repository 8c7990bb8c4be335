"""How the L2 flush between timed iterations is done changes what the next kernel pays for: a memset leaves
126 MB of DIRTY lines whose write-back lands inside the next kernel; a memset followed by a large read leaves
the L2 cold AND clean.  Times the bench kernel under: no flush, memset, memset + read."""
import os
import sys

import ctypes as C

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['MPE_TC_TIMELINE'] = '1'
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200 import _lib  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402

dev = torch.device('cuda:0')
B = 65536
env = m.make_env('simple_spread', num_envs=B, batched=True, seed=1)
actor = m.FusedActor(random_state_dict(10, 5, 1), seed=1)
env.reset()
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rbuf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)


def run(mode, steps=400):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps + 20):
        if mode in ('memset', 'memset+read'):
            wbuf.zero_()
        if mode in ('memset+read', 'read'):
            rbuf.sum()
        if i >= 20:
            ev[i - 20][0].record()
        env.rollout(actor, 1, record=True)
        if i >= 20:
            ev[i - 20][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    buf = (C.c_ulonglong * (148 * 128))()
    _lib.load().mpe_debug_tc_timeline(buf, 148 * 128)
    a = np.array(list(buf), dtype=np.int64).reshape(148, 4, 32)
    g0, g1, c0, c1 = a[:, 2, 0], a[:, 2, 1], a[:, 2, 2], a[:, 2, 5]
    span = (g1.max() - g0.min()) / 1e3
    mhz = ((c1 - c0) / ((g1 - g0) / 1e3)).mean()
    print('   last launch on device: span %.1f us, longest CTA %d cycles, SM clock from clock64/globaltimer %.0f MHz'
          % (span, (c1 - c0).max(), mhz))
    return sum(t) / len(t), t[len(t) // 2], t[0]


for mode in ('none', 'memset', 'read', 'memset+read', 'none'):
    mean, med, mn = run(mode)
    print('%-12s mean %.2f us  median %.2f us  min %.2f us' % (mode, mean * 1e3, med * 1e3, mn * 1e3))
