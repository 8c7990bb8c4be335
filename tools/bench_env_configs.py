"""Env-only step kernel across BASELINE configs 2/3/4: us per launch, algorithmic GB/s, fraction of measured HBM
peak.  bytes per env step = 41N + 8L + 4ND (+ second action index and comm state for simple_reference,
+ the state word of fullobs_collect_treasure)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402


def _gpu_spin(torch, cycles=20_000_000):
    """~10 ms of GPU-side spinning before a timed launch loop, so that the host enqueues ahead of the GPU and the
    kernels run back to back (torch.cuda._sleep is a private helper: skipped quietly where it does not exist)."""
    spin = getattr(torch.cuda, '_sleep', None)
    if spin is not None:
        spin(cycles)

CONFIGS = [('simple_spread', None, 1 << 20), ('simple_spread', 6, 1 << 19), ('simple_spread', 9, 1 << 18),
           ('simple_spread', 12, 1 << 18), ('simple_reference', None, 1 << 20), ('simple_speaker_listener', None, 1 << 20),
           ('fullobs_collect_treasure', None, 1 << 18)]


def run(dev, peak, verbose=False, reps=30):
    out = []
    for scen, n, B in CONFIGS:
        env = m.make_env(scen, n=n, num_envs=B, batched=True, seed=1, device=dev)
        N, L, D = env.n, env.num_landmarks, env.obs_dim
        nbytes = 41 * N + 8 * L + 4 * N * D
        if scen == 'simple_reference':
            nbytes += 4 * N + 4 * N * 10
        if scen == 'fullobs_collect_treasure':
            nbytes += 8  # the per-env state word (types / alive / holding), read and written; treasure moves are rare
        env.reset()
        au = torch.randint(0, 5, (B, N), dtype=torch.int32, device=dev)
        ac = torch.randint(0, 10, (B, N), dtype=torch.int32, device=dev) if env.act_c else None
        bufs = (torch.empty((B, N, D), device=dev), torch.empty((B, N), device=dev),
                torch.empty((B, N), dtype=torch.uint8, device=dev))
        for _ in range(5):
            env.step(au, ac, out=bufs)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _gpu_spin(torch)  # ~10 ms GPU spin: the timed launches below are all queued before the first one starts
        e0.record()
        for _ in range(reps):
            env.step(au, ac, out=bufs)  # working set per launch is > 126 MB L2 for every config
        e1.record()
        torch.cuda.synchronize(dev)
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        gbs = B * nbytes / sec / 1e9
        rec = {'scenario': scen, 'N': N, 'envs': B, 'bytes_per_env_step': nbytes,
               'us_per_launch': round(sec * 1e6, 2), 'GBps': round(gbs, 1),
               'frac_of_measured_hbm': round(gbs / peak, 3), 'G_agent_steps_per_s': round(B * N / sec / 1e9, 2)}
        out.append(rec)
        if verbose:
            print(json.dumps(rec))
        del env
    return out


if __name__ == '__main__':
    peak = 6528.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = json.load(open(p))['hbm_gbs']
    run(torch.device('cuda:0'), peak, verbose=True)
