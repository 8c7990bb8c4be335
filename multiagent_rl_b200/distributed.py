"""Multi-GPU plumbing: env instances shard trivially (no cross-env term anywhere in World.step,
observation or reward), one process per GPU; the only collective is the reduction of episode-return
statistics (experiments/run.py:23-24,55-57,86-88 keeps these as python lists on one process).

Philox streams are keyed by the GLOBAL env id, so a run's trajectories do not depend on the number
of ranks.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world_size):
    """Contiguous block partition of [0, total_envs): -> (env_id_offset, num_envs) of `rank`."""
    total_envs, rank, world_size = int(total_envs), int(rank), int(world_size)
    if not 0 <= rank < world_size:
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    base, rem = divmod(total_envs, world_size)
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world_size, local_rank).  Single-process runs need no process group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend=backend)
    return rank, world, local


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU, BEFORE any pinned host memory is
    allocated: first-touch then places the staging buffers on the GPU's own NUMA node, so that with one process per
    GPU the ranks' host<->device copies do not all cross one socket's memory controller.  Returns the core list (empty
    when NVML or the affinity call is unavailable - the caller carries on unbound)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w, bits in enumerate(mask) for b in range(64) if (int(bits) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - NVML missing / containers without the affinity syscall
        return []


def reduce_return_stats(local_stats, device=None, group=None):
    """All-reduce [sum(ret), sum(ret^2), n_episodes, n_steps(, n_nonfinite_episodes)] (float64) over ranks and
    derive mean / std of the episode return over the finite episodes.  Works on NCCL (cuda tensor) and gloo (cpu)."""
    t = torch.as_tensor(np.asarray(local_stats, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    s = t.cpu().numpy()
    nonfinite = float(s[4]) if len(s) > 4 else 0.0
    n = s[2] - nonfinite
    mean = s[0] / n if n > 0 else float('nan')
    var = max(s[1] / n - mean * mean, 0.0) if n > 0 else float('nan')
    return {'sum': s[0], 'sumsq': s[1], 'episodes': s[2], 'steps': s[3], 'nonfinite_episodes': nonfinite,
            'mean_return': mean,
            'std_return': float(np.sqrt(var)) if n > 0 else float('nan')}


def max_over_ranks(value, device=None, group=None):
    """Device time of a multi-GPU step is the MAX over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_replay(batch, group=None):
    """All-gather a sampled replay shard (same shape on every rank) -> concatenated along dim 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return batch
    out = [torch.empty_like(batch) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, batch.contiguous(), group=group)
    return torch.cat(out, dim=0)
