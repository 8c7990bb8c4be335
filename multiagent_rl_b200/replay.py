"""Device-resident replay ring with the reference's ``ReplayBuffer`` surface, batched over env instances.

Mirrors ``rls/replay_buffer.py:9-91`` (``add`` / ``make_index`` / ``sample_index`` / ``sample`` / ``collect`` /
``clear`` / ``__len__``); what is stored per transition is the tuple of ``experiments/run.py:46,52``:
``(obs_n, action_n_env, rew_shared = np.sum(rew_n), new_obs_n, float(done))``.  Transitions of a batched step are
appended in env order, as if the reference loop had added them one by one.  Everything stays in HBM.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class DeviceReplayBuffer(object):
    def __init__(self, size, num_agents, obs_dim, act_dims, device=None, seed=0):
        if not torch.cuda.is_available():
            raise RuntimeError('multiagent_rl_b200 needs a CUDA device: the replay ring lives in HBM')
        self._lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        act_dims = [int(act_dims)] if not isinstance(act_dims, (list, tuple)) else [int(a) for a in act_dims]
        self.N, self.D, self.act_dims, self.A = int(num_agents), int(obs_dim), act_dims, sum(act_dims)
        self.A1 = act_dims[1] if len(act_dims) > 1 else 0
        self._maxsize = int(size)
        self.seed = int(seed)
        cfg = _lib.ReplayConfig(capacity=self._maxsize, num_agents=self.N, obs_dim=self.D, act0=act_dims[0],
                                act1=act_dims[1] if len(act_dims) > 1 else 0, device=self.device.index, reserved0=0)
        h = C.c_void_p()
        _lib.check(self._lib.replay_create(C.byref(cfg), C.byref(h)), 'replay_create')
        self._h = h

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and self._lib is not None:
            self._lib.replay_destroy(h)
            self._h = None

    def __len__(self):
        return int(self._lib.replay_len(self._h))

    @property
    def _next_idx(self):
        return int(self._lib.replay_next_idx(self._h))

    def clear(self):
        _lib.check(self._lib.replay_clear(self._h), 'replay_clear')

    def _f32(self, x):
        return None if x is None else torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()

    def _i32(self, x):
        return None if x is None else torch.as_tensor(x, dtype=torch.int32, device=self.device).contiguous()

    def add(self, obs_t, action, reward, obs_tp1, done=None, act_c=None):
        """obs_t / obs_tp1 [B,N,D]; action = act_u [B,N] int (movement head index), act_c [B,N] for a two-head
        actor; reward [B,N] per-agent rewards (summed to the shared reward like run.py:46); done [B] or None."""
        obs, nxt, rew = self._f32(obs_t), self._f32(obs_tp1), self._f32(reward)
        au, ac, dn = self._i32(action), self._i32(act_c), self._f32(done)
        if obs.dim() != 3 or tuple(obs.shape[1:]) != (self.N, self.D):
            raise ValueError('obs_t must be [B, %d, %d], got %s' % (self.N, self.D, tuple(obs.shape)))
        B = obs.shape[0]
        if self.A1 > 0 and ac is None:
            raise ValueError('act_c is required for a two-head actor')
        for name, t, want in (('obs_tp1', nxt, B * self.N * self.D), ('action', au, B * self.N),
                              ('act_c', ac, B * self.N), ('reward', rew, B * self.N), ('done', dn, B)):
            if t is not None and t.numel() != want:  # the kernels index these raw pointers by B, N, D
                raise ValueError('%s has %d elements, expected %d' % (name, t.numel(), want))
        _lib.check(self._lib.replay_add(self._h, _lib.ptr(obs), _lib.ptr(au), _lib.ptr(ac), _lib.ptr(rew),
                                        _lib.ptr(nxt), _lib.ptr(dn), B, _lib.current_stream(self.device)), 'replay_add')

    def _gather(self, batch, idx):
        dev = self.device
        out = (torch.empty((batch, self.N, self.D), device=dev), torch.empty((batch, self.N, self.A), device=dev),
               torch.empty((batch,), device=dev), torch.empty((batch, self.N, self.D), device=dev),
               torch.empty((batch,), device=dev))
        idx_out = torch.empty((batch,), dtype=torch.int64, device=dev) if idx is None else None
        _lib.check(self._lib.replay_sample(self._h, batch, _lib.ptr(idx), C.c_uint64(self.seed), _lib.ptr(out[0]),
                                           _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]), _lib.ptr(out[4]),
                                           _lib.ptr(idx_out), _lib.current_stream(dev)), 'replay_sample')
        return out, idx_out

    def make_index(self, batch_size):
        """Uniform indices with replacement (``random.randint(0, len - 1)`` per sample), on the device."""
        batch = int(batch_size)
        idx = torch.empty((batch,), dtype=torch.int64, device=self.device)
        _lib.check(self._lib.replay_sample(self._h, batch, None, C.c_uint64(self.seed), None, None, None, None, None,
                                           _lib.ptr(idx), _lib.current_stream(self.device)), 'replay_sample')
        return idx

    def make_latest_index(self, batch_size):
        idx = (self._next_idx - 1 - torch.arange(int(batch_size), device=self.device)) % self._maxsize
        return idx[torch.randperm(int(batch_size), device=self.device)]

    def sample_index(self, idxes):
        """-> (obs [n,N,D], act one-hot [n,N,A], rew_shared [n], obs_next [n,N,D], done [n]) device tensors."""
        if not (isinstance(idxes, torch.Tensor) and idxes.is_cuda):
            host = np.asarray(idxes.cpu() if isinstance(idxes, torch.Tensor) else idxes, dtype=np.int64).reshape(-1)
            n = len(self)
            if host.size and (host.min() < -n or host.max() >= n):  # list indexing of ReplayBuffer._storage
                raise IndexError('replay index out of range (len %d)' % n)
        # device-resident indices cannot be checked without a sync: the gather kernel wraps negatives and clamps
        idx = torch.as_tensor(idxes, dtype=torch.int64, device=self.device).contiguous()
        out, _ = self._gather(idx.numel(), idx)
        return out

    def sample(self, batch_size):
        if batch_size > 0:
            out, _ = self._gather(int(batch_size), None)
            return out
        return self.sample_index(torch.arange(len(self), device=self.device))

    def collect(self):
        return self.sample(-1)
