"""CPU tests of the MAAC-fork restatement (oracle/maac_ref.py) behind fullobs_collect_treasure:
hand-derived known answers for the fork's engine additions and for the scenario's pick-up / deposit / respawn and
reward rules, the committed trajectory fixture, and the reference's OWN ``make_env`` +
``local_obs_collect_treasure`` (experiments/scenarios.py:95-121,124-192, compiled into oracle/_ref) executed on it."""
import os

import numpy as np
import pytest

from oracle import build_ref, maac_ref, maac_vec, mpe_ref, philox

P = maac_ref.pack_flags


def _env(benchmark=False):
    env = mpe_ref.make_env('fullobs_collect_treasure', benchmark=benchmark)
    env.scenario.draws = maac_ref.PhiloxDraws(11, 5)
    return env


def _far_state():
    """nothing touches anything: collectors on a row at y = 0.8, deposits at y = -0.8, treasures at y = 0"""
    pos = np.array([[-0.9 + 0.3 * i, 0.8] for i in range(6)] + [[-0.5, -0.8], [0.5, -0.8]])
    tr = np.array([[-0.9 + 0.3 * l, 0.0] for l in range(6)])
    return pos, np.zeros((8, 2)), tr


def _acts(idx):
    return [np.eye(5)[k] for k in idx]


def test_world_constants_and_spaces():
    env = _env()
    w = env.world
    assert env.n == 8 and len(w.landmarks) == 6 and w.cache_dists is True
    assert [a.size for a in w.agents] == [0.05] * 6 + [0.075] * 2
    assert [a.mass for a in w.agents] == [1.0] * 6 + [2.25] * 2
    assert all(a.accel == 1.5 and a.max_speed == 1.0 and a.collide and a.silent for a in w.agents)
    assert all(l.size == 0.025 and not l.collide and l.respawn_prob == 1.0 for l in w.landmarks)
    assert [s.shape for s in env.observation_space] == [(30,)] * 8      # uniform rows: main.py:51 builds ONE actor
    assert all(s.n == 5 for s in env.action_space) and env.shared_reward is False


def test_action_force_is_mass_times_accel_squared():
    """_set_action scales by accel, apply_action_force by mass * accel: dv = (m a)(a u) / m * dt = 0.225 for every
    agent, whatever its mass."""
    env = _env()
    pos, vel, tr = _far_state()
    maac_ref.set_state(env, pos, vel, tr, P([0] * 6, [True] * 6, [-1] * 6))
    env.step(_acts([1, 2, 3, 4, 0, 1, 1, 4]))
    v = np.stack([a.state.p_vel for a in env.world.agents])
    want = 0.225 * np.array([[1, 0], [-1, 0], [0, 1], [0, -1], [0, 0], [1, 0], [1, 0], [0, -1]])
    assert np.allclose(v, want, rtol=0, atol=1e-15)
    p = np.stack([a.state.p_pos for a in env.world.agents])
    assert np.allclose(p, pos + 0.1 * want, rtol=0, atol=1e-15)


def test_max_speed_clip():
    env = _env()
    pos, vel, tr = _far_state()
    vel[0] = [3.0, 4.0]          # 0.75 * 5 = 3.75 > 1 -> direction kept, speed 1
    vel[6] = [0.0, -1.2]         # 0.9 < 1 -> untouched
    maac_ref.set_state(env, pos, vel, tr, P([0] * 6, [True] * 6, [-1] * 6))
    env.step(_acts([0] * 8))
    assert np.allclose(env.world.agents[0].state.p_vel, [0.6, 0.8], atol=1e-15)
    assert np.allclose(env.world.agents[6].state.p_vel, [0.0, -0.9], atol=1e-15)


def test_contact_force_mass_ratio():
    """collector (m = 1) against deposit (m = 2.25) at exactly dist_min = 0.125: |f| = 100 k ln 2; the collector
    receives ratio * f, the deposit f / ratio, so dv = 2.25 f dt and f / 2.25 / 2.25 dt."""
    env = _env()
    pos, vel, tr = _far_state()
    pos[6] = [0.0, -0.8]
    pos[0] = [0.125, -0.8]
    maac_ref.set_state(env, pos, vel, tr, P([0] * 6, [True] * 6, [-1] * 6))
    env.step(_acts([0] * 8))
    f = 100 * 1e-3 * np.log(2.0)
    assert np.isclose(env.world.agents[0].state.p_vel[0], 2.25 * f * 0.1, rtol=1e-9)
    assert np.isclose(env.world.agents[6].state.p_vel[0], -(f / 2.25) / 2.25 * 0.1, rtol=1e-9)
    # two collectors: plain equal and opposite forces
    pos, vel, tr = _far_state()
    pos[1] = pos[0] + [0.1, 0.0]
    maac_ref.set_state(env, pos, vel, tr, P([0] * 6, [True] * 6, [-1] * 6))
    env.step(_acts([0] * 8))
    assert np.isclose(env.world.agents[0].state.p_vel[0], -f * 0.1, rtol=1e-9)
    assert env.world.agents[1].state.p_vel[0] == -env.world.agents[0].state.p_vel[0]


def test_pickup_reward_respawn_deposit_cycle():
    env = _env(benchmark=True)
    pos, vel, tr = _far_state()
    pos[2] = tr[4] + [0.06, 0.0]     # collector 2 on treasure 4 (type 1; contact radius 0.075)
    pos[3] = tr[4] + [-0.07, 0.0]    # collector 3 too (0.13 from collector 2: no contact): the lower index takes it
    types = [0, 0, 0, 0, 1, 0]
    maac_ref.set_state(env, pos, vel, tr, P(types, [True] * 6, [-1] * 6))
    env.scenario.draws.tstep = 0
    o, r, d, info = env.step(_acts([0] * 8))
    # both touch treasure 4 while holding nothing: global collecting reward 2 * 5 for everybody
    glob = 10
    assert info['n'] == [0, 0, 1, 1, 0, 0, 0, 0]
    assert np.isclose(r[2], -0.1 * 0.06 + glob) and np.isclose(r[3], -0.1 * 0.07 + glob)
    assert np.isclose(r[0], -0.1 * 0.8 + glob)        # nearest treasure straight below at distance 0.8
    # two collectors that end the step in contact with each other: -5 each (approach at 0.1 -> 0.075 after damping)
    q, qv = pos.copy(), vel.copy()
    q[0] = [0.0, 0.5]; q[1] = [0.104, 0.5]; qv[1] = [-0.1, 0.0]
    maac_ref.set_state(env, q, qv, tr, P(types, [True] * 6, [-1] * 6))
    r2 = env.step(_acts([0] * 8))[1]
    gap = env.world.agents[1].state.p_pos[0] - env.world.agents[0].state.p_pos[0]
    assert 0.096 < gap < 0.0966 and np.isclose(r2[0], -5 - 0.1 * 0.5 + glob) and np.isclose(r2[1], r2[0], atol=0.02)
    assert np.isclose(r2[4], -0.1 * 0.8 + glob)       # the others do not pay for it
    maac_ref.set_state(env, pos, vel, tr, P(types, [True] * 6, [-1] * 6))
    env.scenario.draws.tstep = 0
    env.step(_acts([0] * 8))
    t, a, h = maac_ref.unpack_flags(maac_ref.get_flags(env))
    assert a == [True, True, True, True, False, True] and h == [-1, -1, 1, -1, -1, -1]
    assert np.all(env.world.landmarks[4].state.p_pos == -999.0)
    # next step: the dead treasure is seen 999 away (sorted last), then respawns from the Philox stream
    env.scenario.draws.tstep = 1
    o, r, d, info = env.step(_acts([0] * 8))
    assert np.allclose(o[2][26:28], -999.0 - pos[2]) and o[2][4:6].tolist() == [0.0, 1.0]
    want_pos, want_type = philox.treasure_respawn(11, np.array([5]), 0, 1, 4)
    assert np.array_equal(env.world.landmarks[4].state.p_pos, want_pos[0])
    t, a, h = maac_ref.unpack_flags(maac_ref.get_flags(env))
    assert a == [True] * 6 and t[4] == int(want_type[0])
    # collector 2 (holding type 1) walks into deposit 1 (agent 7): global deposit reward, then it lets go
    pos2 = np.stack([ag.state.p_pos for ag in env.world.agents])
    pos2[2] = pos2[7] + [0.13, 0.0]
    vel2 = np.zeros((8, 2)); vel2[2] = [-0.2, 0.0]    # 0.13 - 0.1 * 0.15 = 0.115 < 0.125 after the step
    trn = np.stack([l.state.p_pos for l in env.world.landmarks])
    maac_ref.set_state(env, pos2, vel2, trn, maac_ref.get_flags(env))
    env.scenario.draws.tstep = 2
    o, r, d, info = env.step(_acts([0] * 8))
    assert info['n'][2] == 1
    dist = np.linalg.norm(env.world.agents[2].state.p_pos - env.world.agents[7].state.p_pos)
    assert dist < 0.125 and np.isclose(r[2], -0.1 * dist + 5)
    assert np.isclose(r[7], -0.1 * dist + 5)          # the deposit is shaped by its nearest matching holder
    assert maac_ref.unpack_flags(maac_ref.get_flags(env))[2][2] == -1
    # the wrong deposit does not take it
    maac_ref.set_state(env, pos2, vel2, trn, P(t, [True] * 6, [-1, -1, 0, -1, -1, -1]))
    env.step(_acts([0] * 8))
    assert maac_ref.unpack_flags(maac_ref.get_flags(env))[2][2] == 0


def test_deposit_reward_without_holder_is_mean_offset_of_the_others():
    env = _env()
    pos, vel, tr = _far_state()
    maac_ref.set_state(env, pos, vel, tr, P([0] * 6, [True] * 6, [-1] * 6))
    o, r, d, info = env.step(_acts([0] * 8))
    p = np.stack([a.state.p_pos for a in env.world.agents])
    for d_i in (6, 7):
        others = [j for j in range(8) if j != d_i]
        assert np.isclose(r[d_i], -0.1 * np.linalg.norm((p[others] - p[d_i]).mean(axis=0)), rtol=1e-12)


def test_observation_layout():
    """experiments/scenarios.py:95-121: [pos, vel, holding one-hot, 6 x (offset, type one-hot) nearest first]"""
    env = _env()
    pos, vel, tr = _far_state()
    vel[1] = [0.1, -0.2]
    maac_ref.set_state(env, pos, vel, tr, P([1, 0, 1, 0, 0, 1], [True] * 6, [0, 1, -1, -1, -1, -1]))
    o = mpe_ref.get_obs(env)
    assert np.array_equal(o[1][:6], [pos[1][0], pos[1][1], 0.1, -0.2, 0.0, 1.0])
    assert o[0][4:6].tolist() == [1.0, 0.0] and o[6][4:6].tolist() == [0.0, 0.0]
    # collector 1 sits above treasure 1; then 0 and 2 tie (0.3 to either side): the lower index first
    order = [1, 0, 2, 3, 4, 5]
    types = [1, 0, 1, 0, 0, 1]
    for k, l in enumerate(order):
        assert np.allclose(o[1][6 + 4 * k:8 + 4 * k], tr[l] - pos[1])
        assert o[1][8 + 4 * k:10 + 4 * k].tolist() == [float(types[l] == 0), float(types[l] == 1)]


def test_philox_treasure_streams():
    a, t, ty = philox.treasure_reset(3, np.arange(2000), 1)
    assert a.shape == (2000, 8, 2) and t.shape == (2000, 6, 2) and ty.shape == (2000, 6)
    assert np.abs(t).max() < 0.95 and np.abs(a).max() < 1.0 and set(np.unique(ty)) == {0, 1}
    assert abs(ty.mean() - 0.5) < 0.03
    p0, _ = philox.treasure_respawn(3, np.arange(4), 0, 0, 2)
    p1, _ = philox.treasure_respawn(3, np.arange(4), 0, 1, 2)
    assert not np.array_equal(p0, p1)


def test_oracle_reproduces_the_committed_treasure_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, 'mpe_fullobs_collect_treasure.npz'))
    env = mpe_ref.make_env('fullobs_collect_treasure')
    T, B = g['act_u'].shape[:2]
    for b in range(0, B, 5):
        env.scenario.draws = maac_ref.PhiloxDraws(int(g['seed']), b)
        maac_ref.set_state(env, g['pos0'][b], g['vel0'][b], g['tr0'][b], g['flags0'][b])
        assert np.array_equal(np.stack(mpe_ref.get_obs(env)), g['obs0'][b])
        for t in range(T):
            env.scenario.draws.tstep = t
            o, r, d, _ = env.step(_acts(g['act_u'][t, b]))
            assert np.array_equal(np.stack(o), g['obs'][t, b]) and np.array_equal(np.array(r), g['rew'][t, b])
            assert maac_ref.get_flags(env) == g['flags'][t, b]


def test_vectorised_oracle_equals_the_loop_oracle(golden_dir):
    """oracle/maac_vec.py (what the full-size GPU tests compare with) on the fixture: observations, positions, treasure
    positions, state word and benchmark flags BIT-equal to the loop oracle's committed output; rewards within one ulp
    (the loop's ``np.linalg.norm`` of the deposit's 1-D mean offset goes through BLAS dot, the vectorised sum does not)."""
    g = np.load(os.path.join(golden_dir, 'mpe_fullobs_collect_treasure.npz'))
    T, B = g['act_u'].shape[:2]
    v = maac_vec.VecTreasure(B, seed=int(g['seed']))
    v.episode[:] = 0
    v.set_state(g['pos0'], g['vel0'], g['tr0'], g['flags0'])
    assert np.array_equal(v.observe(), g['obs0']) and np.array_equal(v.flags(), g['flags0'])
    for t in range(T):
        o, r, info = v.step(g['act_u'][t])
        assert np.array_equal(o, g['obs'][t]) and np.array_equal(v.flags(), g['flags'][t]), t
        assert np.array_equal(v.pos, g['pos'][t]) and np.array_equal(v.vel, g['vel'][t]) and np.array_equal(v.tr, g['tr'][t])
        assert np.array_equal(info, g['info'][t]) and np.abs(r - g['rew'][t]).max() <= 6e-17, t
    # Philox resets: same streams as the kernels (and as PhiloxDraws behind the loop oracle)
    v2 = maac_vec.VecTreasure(5, seed=11, gid0=5)
    o = v2.reset()
    env = _env()
    assert np.array_equal(np.stack(env.reset()), o[0])


@pytest.mark.skipif(not build_ref.available(), reason='oracle/_ref not built (needs /root/reference)')
def test_reference_make_env_runs_collect_treasure_with_its_own_observation(golden_dir):
    """The reference's make_env (scenarios.py:124-192) patches in ITS local_obs_collect_treasure (:95-121), finds the
    post_step hook (:174-177) and hands it to MultiAgentEnv; on the fixture's states and actions its observations and
    rewards equal the committed ones bit for bit.  The stub scenario carries the fork's STOCK observation (other
    agents visible, no holding flag for deposits), so the 30-wide rows can only come from the reference's function."""
    from tests import _refloop
    S, _, _ = _refloop.use_reference('oracle')
    try:
        env = S.make_env('fullobs_collect_treasure', benchmark=False, discrete_action=True, local_observation=True)
        code = env.observation_callback.__func__.__code__
        assert code.co_filename == 'reference/experiments/scenarios.py' and code.co_name == 'local_obs_collect_treasure'
        assert env.post_step_callback is not None and env.post_step_callback.__name__ == 'post_step'
        assert env.shared_reward is False and env.force_discrete_action is True
        assert [s.shape for s in env.observation_space] == [(30,)] * 8 and env.action_space[0].n == 5
        g = np.load(os.path.join(golden_dir, 'mpe_fullobs_collect_treasure.npz'))
        T, B = g['act_u'].shape[:2]
        scen = env.post_step_callback.__self__
        env.scenario = scen
        for b in range(0, B, 3):
            scen.draws = maac_ref.PhiloxDraws(int(g['seed']), b)
            maac_ref.set_state(env, g['pos0'][b], g['vel0'][b], g['tr0'][b], g['flags0'][b])
            assert np.array_equal(np.stack(mpe_ref.get_obs(env)), g['obs0'][b])
            for t in range(T):
                scen.draws.tstep = t
                o, r, d, info = env.step(_acts(g['act_u'][t, b]))
                assert np.array_equal(np.stack(o), g['obs'][t, b]), (b, t)
                assert np.array_equal(np.array(r), g['rew'][t, b]), (b, t)
                assert maac_ref.get_flags(env) == g['flags'][t, b] and d == [False] * 8
        # the scenario's stock observation is NOT what came out (it is 2 + 2 + 2 + 7 * 8 + 6 * 4 = 86 wide for collectors)
        assert len(scen.stock_observation(env.world.agents[0], env.world)) == 86
        # env.seed + reset: same numpy draws as the oracle's own make_env
        mine = mpe_ref.make_env('fullobs_collect_treasure')
        scen.draws = maac_ref.NumpyDraws()
        env.seed(12345678); oa = env.reset()
        mine.seed(12345678); ob = mine.reset()
        assert all(np.array_equal(x, y) for x, y in zip(oa, ob))
    finally:
        _refloop.purge()
