"""Vectorised float64 version of oracle/maac_ref.py (MAAC-fork engine + fullobs_collect_treasure) over B envs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same operations in the same order as the loop oracle, env-parallel:
tests/test_treasure_oracle.py requires BIT equality with ``maac_ref`` on the committed fixture, which is what lets the
GPU tests compare 65,536-env rollouts with it.  Reset / respawn draws are the kernels' Philox streams
(oracle/philox.py), keyed by (seed, global env id, episode, step).  Citations: see oracle/maac_ref.py.
"""
import numpy as np

from . import philox

N, C, L, D = 8, 6, 6, 30
MASS = np.array([1.0] * C + [2.25] * (N - C))
SIZE = np.array([0.05] * C + [0.075] * (N - C))
TSIZE = 0.025
ACCEL, MAX_SPEED = 1.5, 1.0
DT, DAMPING, CONTACT_FORCE, CONTACT_MARGIN = 0.1, 0.25, 1e+2, 1e-3


class VecTreasure(object):
    def __init__(self, B, seed=0, gid0=0):
        self.B, self.seed = int(B), int(seed)
        self.gid = np.arange(self.B, dtype=np.uint64) + np.uint64(gid0)
        self.pos = np.zeros((B, N, 2)); self.vel = np.zeros((B, N, 2)); self.tr = np.zeros((B, L, 2))
        self.types = np.zeros((B, L), dtype=np.int64)
        self.alive = np.ones((B, L), dtype=bool)
        self.hold = -np.ones((B, C), dtype=np.int64)
        self.episode = np.zeros(B, dtype=np.int64) - 1
        self.tstep = np.zeros(B, dtype=np.int64)

    # ------------------------------------------------------------------ state
    def reset(self):
        self.episode += 1
        a, t, ty = philox.treasure_reset(self.seed, self.gid, self.episode)
        self.pos, self.tr, self.types = a.copy(), t.copy(), ty.copy()
        self.vel = np.zeros((self.B, N, 2))
        self.alive[:] = True
        self.hold[:] = -1
        self.tstep[:] = 0
        return self.observe()

    def set_state(self, pos, vel, tr, flags):
        self.pos, self.vel, self.tr = (np.array(x, dtype=np.float64) for x in (pos, vel, tr))
        f = np.asarray(flags, dtype=np.int64)
        self.types = (f[:, None] >> np.arange(L)) & 1
        self.alive = ((f[:, None] >> (6 + np.arange(L))) & 1).astype(bool)
        self.hold = ((f[:, None] >> (12 + 2 * np.arange(C))) & 3) - 1

    def flags(self):
        f = (self.types << np.arange(L)).sum(1) + (self.alive.astype(np.int64) << (6 + np.arange(L))).sum(1)
        return f + ((self.hold + 1) << (12 + 2 * np.arange(C))).sum(1)

    # ------------------------------------------------------------------ distances (World.calculate_distances)
    def _agent_treasure(self):
        """cached_dist_vect[8 + t, i] = -(pos_i - tr_t) and its norm -> [B, N, L, 2], [B, N, L]"""
        vect = -(self.pos[:, :, None, :] - self.tr[:, None, :, :])
        return vect, np.sqrt(vect[..., 0] * vect[..., 0] + vect[..., 1] * vect[..., 1])

    def _agent_agent(self):
        """cached_dist_mag between agents [B, N, N] and cached_dist_vect[j, i] (offset of j seen from i) [B, N, N, 2]"""
        vect = np.zeros((self.B, N, N, 2))
        for a in range(N):
            for b in range(a + 1, N):
                delta = self.pos[:, a] - self.pos[:, b]
                vect[:, a, b] = delta
                vect[:, b, a] = -delta
        return vect, np.sqrt(vect[..., 0] * vect[..., 0] + vect[..., 1] * vect[..., 1])

    def observe(self):
        vect, mag = self._agent_treasure()
        order = np.argsort(mag, axis=2, kind='stable')       # sorted(zip(dist, index)): ties by index
        obs = np.zeros((self.B, N, D))
        obs[:, :, 0:2], obs[:, :, 2:4] = self.pos, self.vel
        obs[:, :C, 4] = self.hold == 0
        obs[:, :C, 5] = self.hold == 1
        sv = np.take_along_axis(vect, order[..., None], axis=2)
        st = np.take_along_axis(np.broadcast_to(self.types[:, None, :], (self.B, N, L)), order, axis=2)
        lst = obs[:, :, 6:].reshape(self.B, N, L, 4)
        lst[..., 0:2] = sv
        lst[..., 2] = st == 0
        lst[..., 3] = st == 1
        return obs

    # ------------------------------------------------------------------ MultiAgentEnv.step
    def step(self, act_u):
        act_u = np.asarray(act_u)
        B = self.B
        onehot = np.eye(5)[act_u]                                         # [B, N, 5]
        u = np.zeros((B, N, 2))
        u[..., 0] += onehot[..., 1] - onehot[..., 2]
        u[..., 1] += onehot[..., 3] - onehot[..., 4]
        u *= ACCEL                                                        # _set_action: sensitivity = accel
        force = (MASS * ACCEL)[None, :, None] * u                         # apply_action_force: (mass * accel) * u
        avect, amag = self._agent_agent()
        with np.errstate(invalid='ignore', divide='ignore'):
            for a in range(N):                                            # apply_environment_force, (a, b) lexicographic
                for b in range(a + 1, N):
                    delta, dist = avect[:, a, b], amag[:, a, b]
                    pen = np.logaddexp(0, -(dist - (SIZE[a] + SIZE[b])) / CONTACT_MARGIN) * CONTACT_MARGIN
                    f = CONTACT_FORCE * delta / dist[:, None] * pen[:, None]
                    ratio = MASS[b] / MASS[a]
                    force[:, a] = ratio * f + force[:, a]
                    force[:, b] = -(1 / ratio) * f + force[:, b]
            self.vel = self.vel * (1 - DAMPING)                           # integrate_state
            self.vel = self.vel + (force / MASS[None, :, None]) * DT
            speed = np.sqrt(np.square(self.vel[..., 0]) + np.square(self.vel[..., 1]))
            clip = speed > MAX_SPEED
            scaled = self.vel / np.where(clip, speed, 1.0)[..., None] * MAX_SPEED
            self.vel = np.where(clip[..., None], scaled, self.vel)
            self.pos = self.pos + self.vel * DT
        obs = self.observe()
        rew, info, touch_t, touch_d = self._rewards()
        self._post_step(touch_t, touch_d)
        self.tstep += 1
        return obs, rew, info

    def _rewards(self):
        B = self.B
        _, tmag = self._agent_treasure()
        avect, amag = self._agent_agent()
        coll, dep = slice(0, C), slice(C, N)
        touch_t = tmag[:, coll, :] < (SIZE[:C, None] + TSIZE)[None]       # [B, C, L]  is_collision(collector, treasure)
        touch_d = amag[:, coll, C:] < (SIZE[:C, None] + SIZE[None, C:])[None]   # [B, C, 2]
        touch_c = amag[:, coll, coll] < (SIZE[:C, None] + SIZE[None, :C])[None]
        free = self.hold < 0
        g_dep = np.zeros(B, dtype=np.int64)
        for d in range(N - C):
            g_dep += 5 * ((self.hold == d) & touch_d[:, :, d]).sum(1)
        g_col = np.zeros(B, dtype=np.int64)
        for l in range(L):
            g_col += 5 * (free & touch_t[:, :, l]).sum(1)
        glob = g_dep + g_col
        rew = np.zeros((B, N))
        info = np.zeros((B, N), dtype=np.int32)
        for i in range(C):
            ncc = touch_c[:, i, :].sum(1) - touch_c[:, i, i]
            r = (0 - 5 * ncc).astype(np.float64)
            near_t = tmag[:, i, :].min(1)
            hd = np.clip(self.hold[:, i], 0, 1)
            near_d = np.take_along_axis(amag[:, i, C:], hd[:, None], axis=1)[:, 0]
            r = r - 0.1 * np.where(free[:, i], near_t, near_d)
            rew[:, i] = r + glob
            at_dep = ~free[:, i] & np.take_along_axis(touch_d[:, i, :], hd[:, None], axis=1)[:, 0]
            info[:, i] = at_dep | (free[:, i] & touch_t[:, i, :].any(1))
        for d in range(N - C):
            a = C + d
            holders = self.hold == d                                      # [B, C]
            any_h = holders.any(1)
            dmin = np.where(holders, amag[:, a, :C], np.inf).min(1)
            others = [j for j in range(N) if j != a]
            mag_o = amag[:, others, a]
            order = np.argsort(mag_o, axis=1, kind='stable')
            vec_o = np.take_along_axis(avect[:, others, a, :], order[..., None], axis=1)   # offsets, nearest first
            mean = vec_o.mean(axis=1)
            m = np.where(any_h, dmin, np.sqrt(mean[:, 0] * mean[:, 0] + mean[:, 1] * mean[:, 1]))
            rew[:, a] = (0 - 0.1 * m) + glob
        return rew, info, touch_t, touch_d

    def _post_step(self, touch_t, touch_d):
        dead_before = ~self.alive.copy()
        for l in range(L):
            cand = touch_t[:, :, l] & (self.hold < 0) & self.alive[:, l][:, None]
            hit = cand.any(1)
            first = cand.argmax(1)
            b = np.nonzero(hit)[0]
            self.hold[b, first[b]] = self.types[b, l]
            self.alive[b, l] = False
            self.tr[b, l] = -999.0
            r = np.nonzero(dead_before[:, l])[0]
            if len(r):
                p, ty = philox.treasure_respawn(self.seed, self.gid[r], self.episode[r], self.tstep[r], l)
                self.tr[r, l] = p
                self.types[r, l] = ty
                self.alive[r, l] = True
        hd = np.clip(self.hold, 0, 1)
        drop = (self.hold >= 0) & np.take_along_axis(touch_d, hd[..., None], axis=2)[..., 0]
        self.hold[drop] = -1
