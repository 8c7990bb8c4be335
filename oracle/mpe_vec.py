"""Batched (vectorised numpy, float64) restatement of the same algorithm as
``oracle/mpe_ref.py``, so that parity tests can check 10^5..10^6 envs in seconds.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every arithmetic expression is
evaluated in the same order and with the same numpy primitives as the loop
oracle (``np.sqrt(np.sum(np.square(.)))``, ``np.logaddexp``, ``f + p_force`` in
(a, b) lexicographic order), so the two agree BIT-EXACTLY; tests/test_oracle.py
asserts that.  The loop oracle follows upstream ``multiagent`` (see its header
for provenance and the "parity unpinned" note); observations follow
/root/reference/experiments/scenarios.py:6-63.
"""
import numpy as np

DT = 0.1
DAMPING = 0.25
CONTACT_FORCE = 1e+2
CONTACT_MARGIN = 1e-3
SENSITIVITY = 5.0

LANDMARK_COLORS = {
    'simple_reference': np.array([[0.75, 0.25, 0.25], [0.25, 0.75, 0.25], [0.25, 0.25, 0.75]]),
    'simple_speaker_listener': np.array([[0.65, 0.15, 0.15], [0.15, 0.65, 0.15], [0.15, 0.15, 0.65]]),
}


class Spec(object):
    """Static description of a scenario (what make_world fixes)."""

    def __init__(self, name, n=None, max_speed=None, accel=None):
        self.name = name
        if name == 'simple_spread':
            self.N = 3 if n is None else int(n)
            self.L = self.N
            self.dim_c = 2
            self.agent_size = [0.15] * self.N
            self.collide = [True] * self.N
            self.movable = [True] * self.N
            self.silent = [True] * self.N
            self.obs_dim = 4 + 2 * self.L
            self.act_u = 5
            self.act_c = 0
        elif name == 'simple_reference':
            self.N, self.L, self.dim_c = 2, 3, 10
            self.agent_size = [0.05] * 2
            self.collide = [False] * 2
            self.movable = [True] * 2
            self.silent = [False] * 2
            self.obs_dim = 2 + 6 + 3 + 10
            self.act_u = 5
            self.act_c = 10
        elif name == 'simple_speaker_listener':
            self.N, self.L, self.dim_c = 2, 3, 3
            self.agent_size = [0.075] * 2
            self.collide = [False] * 2
            self.movable = [False, True]
            self.silent = [False, True]
            self.obs_dim = 2 + 6 + 3
            self.act_u = 5
            self.act_c = 0  # message shares the single width-5 head (ambiguity 3)
        else:
            raise ValueError(name)
        self.max_speed = max_speed
        self.accel = accel


def _dist(d):
    return np.sqrt(np.sum(np.square(d), axis=-1))


class VecEnv(object):
    """State: pos/vel [B,N,2], lm [B,L,2], goal [B,N] (landmark index or -1), comm [B,N,dim_c]."""

    def __init__(self, spec, B):
        self.spec = spec
        self.B = int(B)
        s = spec
        self.pos = np.zeros((B, s.N, 2))
        self.vel = np.zeros((B, s.N, 2))
        self.lm = np.zeros((B, s.L, 2))
        self.goal = -np.ones((B, s.N), dtype=np.int64)
        self.comm = np.zeros((B, s.N, s.dim_c))

    # -- state ---------------------------------------------------------------
    def set_state(self, pos, vel, lm, goal=None):
        self.pos = np.array(pos, dtype=np.float64).reshape(self.B, self.spec.N, 2)
        self.vel = np.array(vel, dtype=np.float64).reshape(self.B, self.spec.N, 2)
        self.lm = np.array(lm, dtype=np.float64).reshape(self.B, self.spec.L, 2)
        if goal is not None:
            self.goal = np.array(goal, dtype=np.int64).reshape(self.B, self.spec.N)
        self.comm = np.zeros((self.B, self.spec.N, self.spec.dim_c))

    # -- MultiAgentEnv._set_action (one-hot / force_discrete_action branch) ----
    def _forces_from_actions(self, act_u):
        s = self.spec
        onehot = np.zeros((self.B, s.N, 5))
        np.put_along_axis(onehot, act_u[..., None].astype(np.int64), 1.0, axis=-1)
        u = np.zeros((self.B, s.N, 2))
        u[..., 0] += onehot[..., 1] - onehot[..., 2]
        u[..., 1] += onehot[..., 3] - onehot[..., 4]
        u *= (s.accel if s.accel is not None else SENSITIVITY)
        return u

    def step(self, act_u, act_c=None, comm_vec=None):
        """act_u [B,N] int in 0..4; act_c [B,N] int message index (one-hot
        message) or comm_vec [B,N,dim_c] float message.  Returns obs, rew, flags."""
        s = self.spec
        B, N = self.B, s.N
        act_u = np.asarray(act_u).reshape(B, N)
        u = self._forces_from_actions(act_u)
        # World.apply_action_force
        p_force = [u[:, i, :].copy() if s.movable[i] else None for i in range(N)]
        # World.apply_environment_force (landmarks never collide in these scenarios)
        for a in range(N):
            for b in range(a + 1, N):
                if not (s.collide[a] and s.collide[b]):
                    continue
                delta = self.pos[:, a] - self.pos[:, b]
                dist = _dist(delta)
                dist_min = s.agent_size[a] + s.agent_size[b]
                k = CONTACT_MARGIN
                with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
                    pen = np.logaddexp(0, -(dist - dist_min) / k) * k
                    force = CONTACT_FORCE * delta / dist[:, None] * pen[:, None]
                if s.movable[a]:
                    p_force[a] = (+force) + (p_force[a] if p_force[a] is not None else 0.0)
                if s.movable[b]:
                    p_force[b] = (-force) + (p_force[b] if p_force[b] is not None else 0.0)
        # World.integrate_state
        for i in range(N):
            if not s.movable[i]:
                continue
            v = self.vel[:, i] * (1 - DAMPING)
            if p_force[i] is not None:
                v = v + (p_force[i] / 1.0) * DT
            if s.max_speed is not None:
                speed = np.sqrt(np.square(v[:, 0]) + np.square(v[:, 1]))
                over = speed > s.max_speed
                with np.errstate(invalid='ignore', divide='ignore'):
                    vc = v / speed[:, None] * s.max_speed
                v = np.where(over[:, None], vc, v)
            self.vel[:, i] = v
            self.pos[:, i] = self.pos[:, i] + v * DT
        # World.update_agent_state
        for i in range(N):
            if s.silent[i]:
                self.comm[:, i] = 0.0
            elif comm_vec is not None:
                self.comm[:, i] = np.asarray(comm_vec, dtype=np.float64)[:, i, :s.dim_c]
            else:
                c = np.zeros((B, s.dim_c))
                idx = np.asarray(act_c).reshape(B, N)[:, i].astype(np.int64)
                if s.name == 'simple_speaker_listener':
                    # width-5 one-hot head cut to dim_c: indices >= dim_c give all zeros
                    ok = idx < s.dim_c
                    c[np.nonzero(ok)[0], idx[ok]] = 1.0
                else:
                    c[np.arange(B), idx] = 1.0
                self.comm[:, i] = c
        return self.observe(), self.reward(), self.flags()

    # -- observations: experiments/scenarios.py:6-63 ---------------------------
    def observe(self):
        s = self.spec
        B = self.B
        out = np.zeros((B, s.N, s.obs_dim))
        rel = self.lm[:, None, :, :] - self.pos[:, :, None, :]  # [B,N,L,2]
        if s.name == 'simple_spread':
            out[:, :, 0:2] = self.vel
            out[:, :, 2:4] = self.pos
            out[:, :, 4:] = rel.reshape(B, s.N, 2 * s.L)
        else:
            colors = LANDMARK_COLORS[s.name]
            out[:, :, 0:2] = self.vel
            out[:, :, 2:8] = rel.reshape(B, s.N, 6)
            gc = np.where((self.goal >= 0)[..., None], colors[np.maximum(self.goal, 0)], 0.0)
            out[:, :, 8:11] = gc
            if s.name == 'simple_reference':
                out[:, 0, 11:] = self.comm[:, 1]
                out[:, 1, 11:] = self.comm[:, 0]
        return out

    # -- rewards ---------------------------------------------------------------
    def _landmark_min_dists(self):
        s = self.spec
        d = np.stack([np.stack([_dist(self.pos[:, a] - self.lm[:, l]) for a in range(s.N)], axis=-1)
                      for l in range(s.L)], axis=1)  # [B,L,N]
        return d.min(axis=-1)  # [B,L]

    def reward(self):
        s = self.spec
        B = self.B
        rew = np.zeros((B, s.N))
        if s.name == 'simple_spread':
            md = self._landmark_min_dists()
            base = np.zeros(B)
            for l in range(s.L):
                base = base - md[:, l]
            for i in range(s.N):
                r = base.copy()
                for a in range(s.N):
                    hit = _dist(self.pos[:, a] - self.pos[:, i]) < (s.agent_size[a] + s.agent_size[i])
                    r = np.where(hit, r - 1, r)
                rew[:, i] = r
        elif s.name == 'simple_reference':
            ar = np.arange(B)
            for i in range(2):
                g = self.goal[:, i]
                gp = self.lm[ar, np.maximum(g, 0)]
                d2 = np.sum(np.square(self.pos[:, 1 - i] - gp), axis=-1)
                rew[:, i] = np.where(g >= 0, -d2, 0.0)
        else:
            ar = np.arange(B)
            gp = self.lm[ar, np.maximum(self.goal[:, 0], 0)]
            d2 = np.sum(np.square(self.pos[:, 1] - gp), axis=-1)
            rew[:, 0] = -d2
            rew[:, 1] = -d2
        return rew

    def flags(self):
        """simple_spread benchmark_data integers: collisions[B,N] (incl. self) and
        occupied_landmarks[B]; zeros for the other scenarios."""
        s = self.spec
        B = self.B
        coll = np.zeros((B, s.N), dtype=np.int32)
        occ = np.zeros(B, dtype=np.int32)
        if s.name == 'simple_spread':
            for i in range(s.N):
                for a in range(s.N):
                    hit = _dist(self.pos[:, a] - self.pos[:, i]) < (s.agent_size[a] + s.agent_size[i])
                    coll[:, i] += hit.astype(np.int32)
            occ = (self._landmark_min_dists() < 0.1).sum(axis=1).astype(np.int32)
        return coll, occ

    def min_dists_sum(self):
        md = self._landmark_min_dists()
        out = np.zeros(self.B)
        for l in range(self.spec.L):
            out = out + md[:, l]
        return out
