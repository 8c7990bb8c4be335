"""Device replay ring: add and sample throughput (HBM streaming / record gathers; bytes = the payload the kernels must move)."""
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

def run(dev, peak, verbose=False):
    N, D, A = 3, 10, 5
    B, cap, batch = 1 << 20, 1 << 22, 1 << 20
    buf = m.DeviceReplayBuffer(cap, N, D, A, device=dev)
    obs = torch.randn(B, N, D, device=dev); nxt = torch.randn(B, N, D, device=dev)
    rew = torch.randn(B, N, device=dev); au = torch.randint(0, 5, (B, N), dtype=torch.int32, device=dev)

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _gpu_spin(torch)  # ~10 ms GPU spin: the timed launches are all queued before the first one starts
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) * 1e-3 / reps

    t_add = timeit(lambda: buf.add(obs, au, rew, nxt))
    add_bytes = B * (2 * N * D * 4 * 2 + N * 4 + N * 4 + 2 * N + 8)  # read + write obs/obs_next, read act/rew, write act/rew/done
    t_s = timeit(lambda: buf.sample(batch))
    s_bytes = batch * (2 * N * D * 4 * 2 + 2 * N + N * A * 4 + 8 + 8 + 8)
    out = [{'op': 'replay_add', 'transitions': B, 'us': round(t_add * 1e6, 1), 'GBps': round(add_bytes / t_add / 1e9, 1),
            'frac_of_measured_hbm': round(add_bytes / t_add / 1e9 / peak, 3), 'G_transitions_per_s': round(B / t_add / 1e9, 2)},
           {'op': 'replay_sample (uniform, random gather of whole records)', 'transitions': batch, 'us': round(t_s * 1e6, 1),
            'GBps': round(s_bytes / t_s / 1e9, 1), 'frac_of_measured_hbm': round(s_bytes / t_s / 1e9 / peak, 3),
            'G_transitions_per_s': round(batch / t_s / 1e9, 2)}]
    if verbose:
        for r in out:
            print(json.dumps(r))
    del buf
    return out


if __name__ == '__main__':
    peak = 6528.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = json.load(open(p))['hbm_gbs']
    run(torch.device('cuda:0'), peak, verbose=True)
