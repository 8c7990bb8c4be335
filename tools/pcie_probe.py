"""What the host-buffer path can reach on this box: pinned-copy bandwidth and per-copy fixed cost in each direction,
alone and full duplex, with and without binding the process to the GPU's NUMA-local cores first; then HostRollout at
1..4 shards.  Prints one JSON object (profiles/r2_pcie_probe.json is a run of this)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200 import distributed as D  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402


def copies(tag, out):
    dev = torch.device('cuda', 0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mb in (0.25, 1, 2, 4, 8, 32):
        n = int(mb * (1 << 20))
        h_in = torch.zeros(n, dtype=torch.uint8).pin_memory()
        h_out = torch.zeros(n, dtype=torch.uint8).pin_memory()
        d_in, d_out = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
        reps = 50

        def run(up, down):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                if up:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e6
        run(True, True)
        res['%g MiB' % mb] = {'h2d_us': run(True, False), 'd2h_us': run(False, True), 'both_us': run(True, True)}
    out[tag] = res


def chunked(out):
    """8 MiB host->device as one copy vs a train of smaller copies on one stream (some boxes show a ~300 us floor
    for 2 - 8 MiB H2D copies while 1 MiB copies run at link speed)."""
    dev = torch.device('cuda', 0)
    n = 8 << 20
    h = torch.zeros(n, dtype=torch.uint8).pin_memory()
    d = torch.zeros(n, dtype=torch.uint8, device=dev)
    h2 = torch.zeros(n, dtype=torch.uint8).pin_memory()
    res = {}
    for kb in (128, 256, 512, 1024, 2048, 8192):
        c = kb << 10
        for direction in ('h2d', 'd2h'):
            def run():
                for o in range(0, n, c):
                    if direction == 'h2d':
                        d[o:o + c].copy_(h[o:o + c], non_blocking=True)
                    else:
                        h2[o:o + c].copy_(d[o:o + c], non_blocking=True)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(30):
                run()
            torch.cuda.synchronize()
            res['%s %d KiB chunks' % (direction, kb)] = (time.perf_counter() - t0) / 30 * 1e6
    out['8MiB_chunked_us'] = res


def rollout(out):
    B = 65536
    actor = m.FusedActor(random_state_dict(10, 5, 1), seed=1)
    res = {}
    for shards in (1, 2, 3, 4, 6):
        torch.cuda.synchronize()
        for depth in (1, 2, 3):
            hr = m.HostRollout('simple_spread', B, actor, shards=shards, seed=1, depth=depth)
            cb = lambda k, tr: tr.rew_np[:64].sum()  # noqa: E731
            for _ in range(10):
                hr.step(cb)
            hr.flush(cb)
            t0 = time.perf_counter()
            for _ in range(100):
                hr.step(cb)
            hr.flush(cb)
            dt = (time.perf_counter() - t0) / 100
            res['%d shards depth %d' % (shards, depth)] = {'us_per_step': dt * 1e6, 'G_agent_steps_per_s': B * 3 / dt / 1e9}
            hr.close()
    out['host_rollout_65536'] = res


if __name__ == '__main__':
    # one mode per process (torch caches pinned blocks, so a later binding would not move them)
    mode = sys.argv[1] if len(sys.argv) > 1 else 'bound'
    out = {'mode': mode, 'cpu_count': os.cpu_count(), 'affinity_before': len(os.sched_getaffinity(0))}
    if mode == 'bound':
        out['numa_cores'] = D.bind_to_gpu_numa(0)
    copies('copies', out)
    chunked(out)
    rollout(out)
    print(json.dumps(out))
