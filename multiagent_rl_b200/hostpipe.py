"""Host-buffer rollouts at PCIe full duplex: the reference's act -> step loop for a caller whose observations,
actions and rewards live in HOST memory (numpy / pinned tensors), split over a few shards that are kept in
different phases of the loop.

The reference moves every step's inputs to the GPU and its results back (``ddpg_gumbel_fix.py:93-94`` H2D of the
observations, ``:100`` D2H of the actions; the env's observations / rewards / dones are host objects).  A batched
caller that keeps that contract pays two transfers per step in OPPOSITE directions, and the blocking
``actor_forward_host`` / ``mpe_step_host`` pair runs them strictly one after the other.  ``HostRollout`` splits the
envs into ``shards`` independent shards (own env handle, own CUDA stream, own page-locked transition blocks) and
drives them with the non-blocking ``mpe_act_step_host_async`` entry point: while shard A's step downloads its
transition block, shard B's observations upload, so both PCIe directions carry data at once.  Every step still
uploads every shard's observations and downloads its actions, next observations, rewards and dones; the only thing
the fused call drops is the bounce of the sampled actions through host memory between act and step.

Philox streams are keyed by the global env id, so the trajectories do not depend on the number of shards.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .env import BatchedMultiAgentEnv


class Transition(object):
    """Host views of one shard-step, the tuple of experiments/run.py:52: ``obs_prev`` [B_s,N,D] float32 (what the actor
    saw), ``act_u`` [B_s,N] int32 (``act_c`` for a two-head actor), ``rew`` [B_s,N], ``obs`` [B_s,N,D] (after the
    step), ``done`` [B_s,N] uint8 - torch tensors over page-locked memory, ``rew_np`` / ``obs_np`` the same memory as
    numpy arrays.  Valid until ``depth`` further steps have been enqueued for the shard."""
    __slots__ = ('shard', 'step', 'obs_prev', 'act_u', 'act_c', 'rew', 'done', 'obs', 'rew_np', 'obs_np')


class _Shard(object):
    __slots__ = ('env', 'stream', 'sp', 'offset', 'num_envs', 'blocks', 'reset_obs', 'events', 'inflight', 'last')


class HostRollout(object):
    """``HostRollout(scenario, num_envs, actor, shards=3, depth=2)``; ``step(on_shard)`` x K; ``flush(on_shard)``.

    Every ``step()`` enqueues one act -> step iteration (experiments/run.py:36-44) for every shard through
    ``mpe_act_step_host_async``: upload of the shard's observations from host memory, actor forward + Gumbel sample,
    env step, ONE download of the transition block {actions, next observations, rewards, dones}.  A shard's steps are
    ordered by its stream (step i uploads the observations step i-1 downloaded), so the host never has to wait before
    enqueueing: up to ``depth`` steps per shard are in flight and ``on_shard(shard_index, transition)`` is called for
    the step that finished ``depth`` iterations earlier - the copy engines always have the next transfer queued.
    ``flush`` delivers the transitions still in flight; ``wait`` just drains the streams.

    ``depth=1`` keeps the host IN the loop: the callback for step i runs before step i + 1 is enqueued, so whatever it
    writes into ``transition.obs`` (normalisation, masking) is what the actor sees next - at the price of one idle gap
    per shard and step (0.75 instead of 0.83 G agent-steps/s at 65,536 envs).  ``depth=2`` (default) is for callers
    that only consume the transitions (replay writes, statistics).
    """

    def __init__(self, scenario_name, num_envs, actor, shards=3, n=None, seed=0, env_id_offset=0, device=None,
                 max_episode_len=25, track_returns=False, depth=2):
        if shards < 1 or depth < 1:
            raise ValueError('shards and depth must be >= 1')
        self.actor = actor
        self.depth = int(depth)
        self.device = actor.device if device is None else torch.device(device)
        lib = self._lib = _lib.load()
        self.max_episode_len = int(max_episode_len)
        self.num_envs = int(num_envs)
        base, rem = divmod(self.num_envs, shards)
        self.shards = []
        off = 0
        for k in range(shards):
            nb = base + (1 if k < rem else 0)
            if nb == 0:
                continue
            sh = _Shard()
            sh.env = BatchedMultiAgentEnv(scenario_name, n=n, num_envs=nb, device=self.device, seed=seed,
                                          env_id_offset=env_id_offset + off, batched=True,
                                          max_episode_len=max_episode_len)
            if track_returns:
                sh.env.track_returns(True)
            sh.stream = torch.cuda.Stream(device=self.device)
            sh.sp = C.c_void_p(sh.stream.cuda_stream)
            sh.offset, sh.num_envs, sh.inflight, sh.last = off, nb, [], None
            sh.events = [torch.cuda.Event() for _ in range(self.depth + 1)]
            self.shards.append(sh)
            off += nb
        probe = self.shards[0].env
        self.n, self.obs_dim, self.two_heads = probe.n, probe.obs_dim, probe.act_c > 0
        if actor.obs_dim != self.obs_dim:
            raise ValueError('actor expects D=%d, env gives %d' % (actor.obs_dim, self.obs_dim))
        N, D, R = self.n, self.obs_dim, self.depth + 1
        # one cudaHostAlloc allocation (NOT tensor.pin_memory(): see _lib.HostBlock) holding, per shard, a ring of
        # depth + 1 transition blocks laid out as mpe_host_block_layout says (a step's results arrive with a single
        # copy) and one buffer for the observations of a fresh episode
        layouts = []
        for sh in self.shards:
            lay = _lib.MpeHostBlockLayout()
            _lib.check(lib.mpe_host_block_layout(sh.env._h, C.byref(lay)), 'mpe_host_block_layout')
            layouts.append(lay)
        up = lambda x: (x + 255) // 256 * 256  # noqa: E731
        total = sum(R * int(l.bytes) + up(sh.num_envs * N * D * 4) for sh, l in zip(self.shards, layouts))
        self._block = _lib.HostBlock(total + 256)
        pos = 0
        for sh, lay in zip(self.shards, layouts):
            nb = sh.num_envs
            sh.blocks = []
            for _ in range(R):
                t = lambda shape, dt, o: self._block.tensor(shape, dt, offset=pos + int(o))  # noqa: E731
                b = {'ptr': C.c_void_p(self._block.ptr + pos),
                     'act_u': t((nb, N), torch.int32, lay.off_act_u),
                     'act_c': t((nb, N), torch.int32, lay.off_act_c) if self.two_heads else None,
                     'obs': t((nb, N, D), torch.float32, lay.off_obs),
                     'rew': t((nb, N), torch.float32, lay.off_rew),
                     'done': t((nb, N), torch.uint8, lay.off_done)}
                b['obs_ptr'] = C.c_void_p(b['obs'].data_ptr())
                b['rew_np'], b['obs_np'] = b['rew'].numpy(), b['obs'].numpy()
                sh.blocks.append(b)
                pos += int(lay.bytes)
            obs0 = self._block.tensor((nb, N, D), torch.float32, offset=pos)
            sh.reset_obs = {'obs': obs0, 'obs_ptr': C.c_void_p(obs0.data_ptr())}
            pos += up(nb * N * D * 4)
        self.global_step = 0
        self.episode_step = 0
        self._reset_due = True

    # ------------------------------------------------------------------ plumbing
    def close(self):
        """Wait for everything in flight and release the staging block (its tensors become invalid)."""
        if getattr(self, '_block', None) is not None:
            try:
                self.wait()
            except Exception:  # noqa: BLE001 - interpreter shutdown / a failed stream: still release the memory
                pass
            finally:
                self._block.free()
                self._block = None

    __del__ = close

    def wait(self):
        """Block until every shard's enqueued work has finished (no callbacks; see ``flush``)."""
        for sh in self.shards:
            _lib.check(self._lib.mpe_host_wait(sh.sp), 'mpe_host_wait')

    def _transition(self, sh, k, rec):
        step, src, dst = rec
        tr = Transition()
        tr.shard, tr.step = k, step
        tr.obs_prev = src['obs']
        tr.act_u, tr.act_c, tr.rew, tr.done, tr.obs = dst['act_u'], dst['act_c'], dst['rew'], dst['done'], dst['obs']
        tr.rew_np, tr.obs_np = dst['rew_np'], dst['obs_np']
        return tr

    def flush(self, on_shard=None):
        """Deliver the transitions still in flight (oldest first) and leave the streams idle."""
        more = True
        while more:
            more = False
            for k, sh in enumerate(self.shards):
                if sh.inflight:
                    rec, ev = sh.inflight.pop(0)
                    ev.synchronize()
                    if on_shard is not None:
                        on_shard(k, self._transition(sh, k, rec))
                    more = more or bool(sh.inflight)
        self.wait()

    def _cat(self, name):
        """[B, ...] COPY assembled from every shard's LATEST transition (tests, small batches)."""
        parts = []
        for k, sh in enumerate(self.shards):
            parts.append(getattr(self._transition(sh, k, sh.last), name))
        return None if parts[0] is None else torch.cat(parts, 0)

    obs = property(lambda self: self._cat('obs'))
    obs_prev = property(lambda self: self._cat('obs_prev'))
    rew = property(lambda self: self._cat('rew'))
    done = property(lambda self: self._cat('done'))
    act_u = property(lambda self: self._cat('act_u'))
    act_c = property(lambda self: self._cat('act_c'))

    # ------------------------------------------------------------------ the loop of experiments/run.py:28-65
    def step(self, on_shard=None):
        """Enqueue one act -> step iteration of every shard; before that, per shard, wait for the step enqueued
        ``depth`` iterations ago and hand it to ``on_shard(shard_index, transition)`` (the caller's per-step host work:
        replay writes, reward sums).  A new episode starts (env.reset(), experiments/run.py:50,59-60) when the last one
        reached ``max_episode_len`` steps.  Returns without waiting for the work enqueued now."""
        lib = self._lib
        call = lib.mpe_act_step_host_async
        terminal = self._reset_due or (self.max_episode_len > 0 and self.episode_step >= self.max_episode_len)
        i = self.global_step
        step = C.c_uint64(i)
        R = self.depth + 1
        for k, sh in enumerate(self.shards):
            if len(sh.inflight) >= self.depth:
                rec, ev = sh.inflight.pop(0)
                ev.synchronize()
                if on_shard is not None:
                    on_shard(k, self._transition(sh, k, rec))
            dst = sh.blocks[i % R]
            if terminal:
                src = sh.reset_obs
                if lib.mpe_reset_host_async(sh.env._h, src['obs_ptr'], sh.sp) != 0:
                    _lib.check(-2, 'mpe_reset_host_async')
            else:
                src = sh.blocks[(i - 1) % R]
            rc = call(sh.env._h, self.actor._h, src['obs_ptr'], step, dst['ptr'], sh.sp)
            if rc != 0:
                _lib.check(rc, 'mpe_act_step_host_async')
            ev = sh.events[i % R]
            ev.record(sh.stream)
            sh.last = (i, src, dst)
            sh.inflight.append((sh.last, ev))
        if terminal:
            self.episode_step = 0
            self._reset_due = False
        self.global_step += 1
        self.episode_step += 1

    def reset(self):
        """Start a new episode with the next ``step()`` (env.reset() is enqueued there, in stream order)."""
        self._reset_due = True

    def read_stats(self):
        self.wait()
        return np.sum([sh.env.read_stats() for sh in self.shards], axis=0)
