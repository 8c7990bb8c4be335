"""GPU parity tests of the env kernels through the C ABI, against the CPU oracle.

Tolerances (stated once, used everywhere):
  fp64 validation build vs the float64 oracle: 1e-12 absolute on positions/velocities/obs/rewards
       (the only non-identical operations are exp/log1p), integer channels bit-exact.
  fp32 build vs the float64 oracle over a 25-step episode: 5e-5 absolute worst case and 1e-5 at
       the 99.99th percentile on positions/obs, 2e-4 on rewards (sums of N+L terms);
       collision / occupancy flags may differ only when the deciding distance is within 1e-6 of
       its threshold.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mpe_ref, mpe_vec, philox

pytestmark = pytest.mark.gpu

F64_ATOL = 1e-12
F32_ATOL, F32_P9999, F32_REW = 5e-5, 1e-5, 2e-4

FILES = [('simple_spread', None), ('simple_spread', 6), ('simple_spread', 9), ('simple_spread', 12),
         ('simple_reference', None), ('simple_speaker_listener', None)]


def _gold(golden_dir, scenario, n):
    return np.load(os.path.join(golden_dir, 'mpe_%s%s.npz' % (scenario, '' if n is None else '_n%d' % n)))


def _mk(scenario, n, B, precision, **kw):
    import multiagent_rl_b200 as m
    return m.make_env(scenario, n=n, num_envs=B, batched=True, precision=precision, **kw)


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _check(a, b, precision, what, atol32=F32_ATOL):
    err = np.abs(a - b)
    if precision == 'fp64':
        assert err.max() <= F64_ATOL, (what, err.max())
    else:
        assert err.max() <= atol32, (what, err.max())


def _flags_differ_only_at_the_threshold(ii, g, t, pos32):
    """fp32 flags vs the float64 fixture: a collision / occupancy flag is a threshold on a distance, so it may differ
    only where the deciding float64 distance lies within the band the fp32 state error of THAT env explains
    (2 x its largest position deviation at this step, + 1e-6 for the rounding of the comparison itself)."""
    N = pos32.shape[1]
    p64, lm = g['pos'][t], g['lm0']
    err = np.abs(pos32 - p64).max(axis=(1, 2))
    band = 2.0 * np.sqrt(2.0) * err + 1e-6
    d = np.linalg.norm(p64[:, :, None] - p64[:, None, :], axis=-1)
    bad = ii[:, :N] != g['coll'][t]
    for b, i in zip(*np.nonzero(bad)):
        assert np.abs(d[b, i] - 0.3).min() <= band[b], ('collision flag', t, b, i, np.abs(d[b, i] - 0.3).min(), band[b])
    dl = np.linalg.norm(p64[:, :, None] - lm[:, None, :], axis=-1).min(axis=1)   # [B, L] nearest agent per landmark
    for b in np.nonzero(ii[:, N] != g['occ'][t])[0]:
        assert np.abs(dl[b] - 0.1).min() <= band[b], ('occupied flag', t, b)
    assert bad.mean() < 0.01


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('scenario,n', FILES)
def test_golden_trajectories(golden_dir, scenario, n, precision):
    g = _gold(golden_dir, scenario, n)
    B = g['pos0'].shape[0]
    env = _mk(scenario, n, B, precision)
    env.set_state(g['pos0'], g['vel0'], g['lm0'], g['goal0'])
    _check(_np(env.observe()), g['obs0'], precision, 'obs0', 1e-6)
    spread = scenario == 'simple_spread'
    for t in range(g['act_u'].shape[0]):
        obs, rew, done, info = env.step(g['act_u'][t], g['act_c'][t] if env.act_c > 0 else None, info=True)
        pos, vel, _, _ = env.get_state()
        _check(_np(pos), g['pos'][t], precision, ('pos', t))
        _check(_np(vel), g['vel'][t], precision, ('vel', t), 5e-4)
        _check(_np(obs), g['obs'][t], precision, ('obs', t), 5e-4)
        _check(_np(rew), g['rew'][t], precision, ('rew', t), F32_REW)
        assert int(done.sum()) == 0 and done.dtype == torch.uint8
        if spread:
            ii = info['info_i'].cpu().numpy()
            if precision == 'fp64':
                assert np.array_equal(ii[:, :-1], g['coll'][t]) and np.array_equal(ii[:, -1], g['occ'][t])
            else:
                _flags_differ_only_at_the_threshold(ii, g, t, _np(pos))
            _check(_np(info['info_f']), -(g['rew'][t][:, 0] + g['coll'][t][:, 0]), precision, 'min_dists', F32_REW)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('scenario,n', [('simple_spread', None), ('simple_spread', 6), ('simple_spread', 9),
                                        ('simple_spread', 12), ('simple_reference', None),
                                        ('simple_speaker_listener', None)])
def test_philox_reset_is_bit_exact(scenario, n, precision):
    B, off, seed = 1000, 12345, 12345678
    env = _mk(scenario, n, B, precision, seed=seed, env_id_offset=off)
    spec = mpe_vec.Spec(scenario, n)
    gid = np.arange(off, off + B)
    for episode in range(3):
        obs = env.reset()
        pos, vel, lm, goal = env.get_state()
        want = philox.reset_positions(seed, gid, episode, spec.N + spec.L)
        assert np.array_equal(_np(pos), want[:, :spec.N]) and np.array_equal(_np(lm), want[:, spec.N:])
        assert float(vel.abs().max()) == 0.0
        v = mpe_vec.VecEnv(spec, B)
        gl = -np.ones((B, spec.N), dtype=np.int64)
        if scenario == 'simple_reference':
            gl = philox.reset_goals(seed, gid, episode, 2, 3)
        elif scenario == 'simple_speaker_listener':
            gl[:, 0] = philox.reset_goals(seed, gid, episode, 1, 3)[:, 0]
        assert np.array_equal(goal.cpu().numpy(), gl)
        v.set_state(want[:, :spec.N], np.zeros((B, spec.N, 2)), want[:, spec.N:], gl)
        _check(_np(obs), v.observe(), precision, 'reset obs', 1e-6)
    # masked reset only touches the selected envs
    before = [x.clone() for x in env.get_state()]
    mask = torch.zeros(B, dtype=torch.uint8)
    mask[::3] = 1
    env.reset(mask=mask)
    after = env.get_state()
    keep = (mask == 0).numpy()
    assert torch.equal(before[0][keep], after[0][keep]) and not torch.equal(before[0][~keep], after[0][~keep])


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_random_rollout_vs_vectorised_oracle(precision):
    """200k envs x 25 steps of random actions from Philox resets (tail warp included: B % 32 != 0)."""
    B, T, seed = 200_003, 25, 7
    env = _mk('simple_spread', None, B, precision, seed=seed)
    env.reset()
    pos, vel, lm, _ = env.get_state()
    v = mpe_vec.VecEnv(mpe_vec.Spec('simple_spread'), B)
    v.set_state(_np(pos), _np(vel), _np(lm))
    rng = np.random.RandomState(0)
    worst = 0.0
    flag_mismatch = 0
    for t in range(T):
        act = rng.randint(0, 5, (B, 3)).astype(np.int32)
        obs, rew, done, info = env.step(act, info=True)
        o, r, (coll, occ) = v.step(act)
        err = np.abs(_np(obs) - o)
        worst = max(worst, err.max())
        ii = info['info_i'].cpu().numpy()
        if precision == 'fp64':
            assert err.max() <= F64_ATOL and np.abs(_np(rew) - r).max() <= F64_ATOL
            assert np.array_equal(ii[:, :3], coll) and np.array_equal(ii[:, 3], occ)
        else:
            assert np.abs(_np(rew) - r)[np.isfinite(r)].max() <= 1.0 + F32_REW  # a flipped flag moves rew by 1
            bad = (ii[:, :3] != coll)
            flag_mismatch += int(bad.sum())
            if bad.any():  # only inside the rounding band of the 0.30 threshold
                d = np.linalg.norm(v.pos[:, :, None] - v.pos[:, None, :], axis=-1)
                near = (np.abs(d - 0.3) < 1e-6).any(axis=(1, 2))
                assert np.all(near[bad.any(axis=1)])
    if precision == 'fp32':
        pos, _, _, _ = env.get_state()
        e = np.abs(_np(pos) - v.pos).max(axis=(1, 2))
        assert e.max() <= F32_ATOL and np.quantile(e, 0.9999) <= F32_P9999, (e.max(), np.quantile(e, 0.9999))
        assert flag_mismatch <= 20


def test_one_step_known_answers_fp32():
    """Hand-derived contact forces: |f| = 100*k*ln2 at contact, ~1.0 at 1 cm overlap, 0 far away."""
    env = _mk('simple_spread', 2, 3, 'fp32')
    pos = np.array([[[0, 0], [0.3, 0]], [[0, 0], [0.29, 0]], [[0, 0], [1.2, 0]]], dtype=np.float64)
    env.set_state(pos, np.zeros((3, 2, 2)), np.zeros((3, 2, 2)) + 5.0)
    env.step(np.zeros((3, 2), dtype=np.int32))
    _, vel, _, _ = env.get_state()
    v = _np(vel)
    f = -v[:, 0, 0] / 0.1  # vel = f * dt
    assert abs(f[0] - 100 * 1e-3 * np.log(2.0)) < 3e-4  # fp32: 0.3f + rounding moves x = d/k by ~3e-5/1e-3
    assert abs(f[1] - 1.0) < 2e-3 and f[2] == 0.0
    assert np.allclose(v[:, 0], -v[:, 1]) and np.all(v[:, :, 1] == 0)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('n', [None, 6])
def test_max_speed_and_accel(precision, n):
    """Entity.max_speed clipping (integrate_state) and Entity.accel (_set_action sensitivity), both builds and both
    kernel families (thread-per-env, lanes-per-env), over 5 steps; about half of the agents are clipped each step."""
    B, N = 4096, n or 3
    rng = np.random.RandomState(1)
    pos, vel, lm = rng.uniform(-1, 1, (B, N, 2)), rng.uniform(-1, 1, (B, N, 2)), rng.uniform(-1, 1, (B, N, 2))
    spec = mpe_vec.Spec('simple_spread', n, max_speed=0.3, accel=3.0)
    v = mpe_vec.VecEnv(spec, B); v.set_state(pos, vel, lm)
    env = _mk('simple_spread', n, B, precision, max_speed=0.3, accel=3.0)
    env.set_state(pos, vel, lm)
    clipped = 0
    for t in range(5):
        act = rng.randint(0, 5, (B, N))
        o, r, _ = v.step(act)
        obs, rew, _, _ = env.step(act)
        speed = np.linalg.norm(v.vel, axis=-1)
        clipped += int((np.abs(speed - 0.3) < 1e-9).sum())
        assert speed.max() <= 0.3 + 1e-12
        if precision == 'fp64':
            assert np.abs(_np(obs) - o).max() <= F64_ATOL and np.abs(_np(rew) - r).max() <= F64_ATOL
        else:
            assert np.abs(_np(obs) - o).max() <= F32_ATOL and np.abs(_np(rew) - r).max() <= 1.0 + F32_REW
            assert np.quantile(np.abs(_np(rew) - r), 0.999) <= F32_REW * N
            _, v32, _, _ = env.get_state()
            assert float(torch.linalg.norm(v32, dim=-1).max()) <= 0.3 * (1 + 2e-7)
    assert clipped > B


def test_list_surface_is_a_drop_in_for_the_reference_loop():
    """experiments/run.py:28-65 shaped loop on num_envs=1 against the loop oracle, same numpy seed."""
    import multiagent_rl_b200 as m
    for scenario in ('simple_spread', 'simple_reference', 'simple_speaker_listener'):
        env = m.make_env(scenario, benchmark=False, discrete_action=True, local_observation=True, precision='fp64')
        ref = mpe_ref.make_env(scenario)
        assert env.n == ref.n and env.observation_space[0].shape[0] == ref.observation_space[0].shape[0]
        if hasattr(ref.action_space[0], 'high'):
            assert (env.action_space[0].high + 1).tolist() == (ref.action_space[0].high + 1).tolist()
            width = int((ref.action_space[0].high + 1).sum())
        else:
            assert env.action_space[0].n == ref.action_space[0].n
            width = ref.action_space[0].n
        env.seed(12345678); obs_n = env.reset()
        ref.seed(12345678); ref_obs = ref.reset()
        rng = np.random.RandomState(0)
        episode_step = 0
        for step in range(60):
            assert isinstance(obs_n, list) and len(obs_n) == env.n and obs_n[0].dtype == np.float64
            for a, b in zip(obs_n, ref_obs):
                assert np.abs(a - b).max() <= F64_ATOL
            logits = rng.randn(env.n, width)
            action_n = [np.array(x) for x in logits.tolist()]
            ref_action = [x.copy() for x in action_n]
            obs_n, rew_n, done_n, info_n = env.step(action_n)
            ref_obs, ref_rew, ref_done, ref_info = ref.step(ref_action)
            for a, b in zip(action_n, ref_action):  # in-place one-hot side effect of _set_action
                assert np.array_equal(a, b)
            assert np.abs(np.array(rew_n) - np.array(ref_rew)).max() <= F64_ATOL
            assert done_n == ref_done and info_n == ref_info
            assert np.isfinite(np.sum(rew_n)) and all(done_n) is False
            episode_step += 1
            if episode_step >= 25:  # both sides draw from numpy's GLOBAL stream: replay the same state
                st = np.random.get_state()
                obs_n = env.reset()
                np.random.set_state(st)
                ref_obs, episode_step = ref.reset(), 0


def test_benchmark_info_tuple():
    import multiagent_rl_b200 as m
    env = m.make_env('simple_spread', benchmark=True, precision='fp64')
    ref = mpe_ref.make_env('simple_spread', benchmark=True)
    env.seed(5); env.reset(); ref.seed(5); ref.reset()
    a = [np.eye(5)[i % 5] for i in range(3)]
    _, _, _, info = env.step([x.copy() for x in a])
    _, _, _, rinfo = ref.step([x.copy() for x in a])
    for x, y in zip(info['n'], rinfo['n']):
        assert abs(x[0] - y[0]) <= F64_ATOL and x[1] == y[1] and abs(x[2] - y[2]) <= F64_ATOL and x[3] == y[3]


def test_discrete_action_input_branch():
    """upstream's index branch (rls/arglist.py:31-36): 1 left, 2 right, 3 down, 4 up, u = +-1 * 5."""
    env = _mk('simple_spread', None, 5, 'fp64')
    env.discrete_action_input = True
    env.set_state(np.zeros((5, 3, 2)) + np.arange(3)[None, :, None], np.zeros((5, 3, 2)), np.zeros((5, 3, 2)) + 9)
    act = np.tile(np.arange(5)[:, None], (1, 3))
    env.step(act)
    _, vel, _, _ = env.get_state()
    v = _np(vel)[:, 0]
    want = np.array([[0, 0], [-0.5, 0], [0.5, 0], [0, -0.5], [0, 0.5]])
    assert np.allclose(v, want, atol=1e-12)


def test_host_buffer_step_matches_device_step():
    B = 4099
    env_a = _mk('simple_spread', None, B, 'fp32', seed=3)
    env_b = _mk('simple_spread', None, B, 'fp32', seed=3)
    env_a.reset(); env_b.reset()
    act = np.random.RandomState(2).randint(0, 5, (B, 3)).astype(np.int32)
    obs, rew, done, _ = env_a.step(act)
    ho, hr, hd = env_b.step_host(act)
    assert np.array_equal(obs.cpu().numpy(), ho) and np.array_equal(rew.cpu().numpy(), hr) and hd.sum() == 0


def test_sharding_is_invariant_to_the_number_of_ranks():
    """Concatenated shard outputs == single-shard outputs for the same global env ids / seed."""
    from multiagent_rl_b200.distributed import shard_range
    B, seed = 10_000, 99
    whole = _mk('simple_spread', None, B, 'fp32', seed=seed)
    o_whole = whole.reset()
    act = np.random.RandomState(4).randint(0, 5, (B, 3)).astype(np.int32)
    s_whole = whole.step(act)
    for world in (2, 3):
        obs, rew = [], []
        for r in range(world):
            off, n = shard_range(B, r, world)
            e = _mk('simple_spread', None, n, 'fp32', seed=seed, env_id_offset=off)
            obs.append(e.reset())
            rew.append(e.step(act[off:off + n])[1])
        assert torch.equal(torch.cat(obs), o_whole) and torch.equal(torch.cat(rew), s_whole[1])


def test_full_size_properties_1m_envs():
    """BASELINE size (1,048,576 envs): size-independent invariants instead of the oracle."""
    B = 1 << 20
    env = _mk('simple_spread', None, B, 'fp32', seed=11)
    env.reset()
    pos0, vel0, lm0, _ = env.get_state()
    act = torch.randint(0, 5, (B, 3), dtype=torch.int32, device='cuda')
    obs, rew, done, _ = env.step(act)
    pos, vel, _, _ = env.get_state()
    # obs layout: [vel, pos, lm - pos]
    assert torch.equal(obs[:, :, 0:2], vel) and torch.equal(obs[:, :, 2:4], pos)
    rel = (lm0[:, None, :, :] - pos[:, :, None, :]).reshape(B, 3, 6)
    assert torch.equal(obs[:, :, 4:], rel)
    # momentum: sum_i vel_i = 0.75 * sum vel0 + 0.1 * sum u  (pair forces cancel) up to fp32 rounding
    table = torch.tensor([[0, 0], [5, 0], [-5, 0], [0, 5], [0, -5]], dtype=torch.float32, device='cuda')
    u = table[act.long()].sum(1)
    assert float((vel.sum(1) - 0.1 * u).abs().max()) < 1e-4
    # reward <= -1 (self collision) and finite; idempotent observe
    assert float(rew.max()) <= -1.0 and bool(torch.isfinite(rew).all())
    assert torch.equal(env.observe(), obs)
    # translation invariance of the relative part: shift everything by a power of two
    env2 = _mk('simple_spread', None, B, 'fp32', seed=11)
    env2.set_state(pos0 + 0.5, vel0, lm0 + 0.5)
    obs2, rew2, _, _ = env2.step(act)
    assert float((obs2[:, :, 4:] - obs[:, :, 4:]).abs().max()) < 1e-5
    assert float((rew2 - rew).abs().max()) < 1e-4


def test_return_tracking_and_stats():
    B = 5000
    env = _mk('simple_spread', None, B, 'fp32', seed=21)
    env.track_returns(True)
    env.reset()
    tot = torch.zeros(B, device='cuda', dtype=torch.float64)
    for t in range(25):
        _, rew, _, _ = env.step(torch.randint(0, 5, (B, 3), dtype=torch.int32, device='cuda'))
        tot += rew.double().sum(1)
    env.reset()
    alias = env.stats_tensor()
    s = env.read_stats(clear=True)
    assert s[2] == B and s[3] == 25 * B
    assert float(alias[2]) == 0.0  # zero-copy view of the (now cleared) device statistics
    assert abs(s[0] - float(tot.sum())) < 1e-2 * B and abs(s[1] - float((tot ** 2).sum())) < 1e-1 * B
    assert env.read_stats()[2] == 0


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_coincident_agents_give_nan_like_upstream(precision):
    """Upstream has no epsilon on the pair distance: two agents at the same point get 0/0 = NaN forces.
    The kernels must reproduce that (same NaN pattern in state, observations and rewards as the loop oracle, which
    keeps Python's `min` semantics for NaN distances) and must not contaminate other envs of the batch."""
    import warnings
    B = 64  # env 5 has agents 0 and 1 coincident; every other env is regular
    rng = np.random.RandomState(4)
    pos = rng.uniform(-1, 1, (B, 3, 2)); vel = rng.uniform(-0.2, 0.2, (B, 3, 2)); lm = rng.uniform(-1, 1, (B, 3, 2))
    pos[5, 1] = pos[5, 0]
    act = rng.randint(0, 5, (B, 3))
    env = _mk('simple_spread', None, B, precision, seed=1)
    env.reset()
    env.set_state(torch.from_numpy(pos), torch.from_numpy(vel), torch.from_numpy(lm))
    obs, rew, _, _ = env.step(torch.from_numpy(act.astype(np.int32)))
    obs, rew = _np(obs), _np(rew)
    o = mpe_ref.make_env('simple_spread')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for b in (4, 5, 6):
            mpe_ref.set_state(o, pos[b], vel[b], lm[b])
            on, rn, _, _ = o.step([np.eye(5)[a] for a in act[b]])
            on, rn = np.stack(on), np.array(rn, dtype=np.float64)
            assert np.array_equal(np.isnan(on), np.isnan(obs[b])), b
            assert np.array_equal(np.isnan(rn), np.isnan(rew[b])), b
            tol = 1e-12 if precision == 'fp64' else 5e-5
            assert np.nanmax(np.abs(on - obs[b])) <= tol and (np.all(np.isnan(rn)) or np.nanmax(np.abs(rn - rew[b])) <= 10 * tol)
    assert np.isnan(obs[5]).any() and not np.isnan(np.delete(obs, 5, 0)).any() and not np.isnan(np.delete(rew, 5, 0)).any()


def test_nonfinite_episodes_are_counted_not_folded():
    """'replicate, don't fix, but count NaNs'.  An env whose agents coincide gets a NaN return; it is
    counted in stats[4] and the finite episodes' sums stay finite."""
    B = 256
    env = _mk('simple_spread', None, B, 'fp32', seed=8)
    env.track_returns(True)
    env.reset()
    pos, vel, lm, _ = env.get_state()
    pos[7, 1] = pos[7, 0]
    pos[100, 2] = pos[100, 0]
    env.set_state(pos, vel, lm)
    tot = torch.zeros(B, device='cuda', dtype=torch.float64)
    for t in range(3):
        _, rew, _, _ = env.step(torch.zeros((B, 3), dtype=torch.int32, device='cuda'))
        tot += rew.double().sum(1)
    env.reset()
    s = env.read_stats()
    assert len(s) == 5 and s[2] == B and s[3] == 3 * B and s[4] == 2
    good = torch.isfinite(tot)
    assert int((~good).sum()) == 2 and abs(s[0] - float(tot[good].sum())) < 1e-3 * B and np.isfinite(s[1])
    from multiagent_rl_b200.distributed import reduce_return_stats
    out = reduce_return_stats(s)
    assert abs(out['mean_return'] - float(tot[good].mean())) < 1e-4


def test_seed_restarts_the_episode_streams_on_the_callers_stream():
    """env.seed(s) (main.py:45): the same seed gives the same episodes again, on a non-default stream too."""
    B = 3000
    env = _mk('simple_spread', None, B, 'fp32', seed=1)
    env.track_returns(True)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        env.seed(77)
        o1 = env.reset().clone()
        env.step(torch.ones((B, 3), dtype=torch.int32, device='cuda'))
        o2 = env.reset().clone()
        env.seed(77)
        o3 = env.reset().clone()
        side.synchronize()
        assert torch.equal(o1, o3) and not torch.equal(o1, o2)
        # the half-finished episode of the old seed was dropped, not folded
        env.reset()
        s = env.read_stats()
    assert s[2] == B and s[3] == B  # only the one-step episode that ended before the re-seed


def test_episode_history_for_env():
    """EpisodeHistory.for_env(env) reads env.max_episode_len."""
    import multiagent_rl_b200 as m
    env = _mk('simple_spread', None, 64, 'fp32', max_episode_len=5)
    h = m.EpisodeHistory.for_env(env)
    assert (h.B, h.N, h.L) == (64, 3, 5)
    env.reset()
    for t in range(5):
        _, rew, _, _ = env.step(torch.zeros((64, 3), dtype=torch.int32, device='cuda'))
        h.add_step(rew)
    assert len(h.finished) == 1 and len(h.history()['reward_episodes']) == 65


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('n', [1, 2, 4, 5, 7, 8, 10, 11])
def test_every_team_size_make_env_accepts(n, precision):
    """experiments/scenarios.py:169-170 passes any ``n`` to make_world(num_agents=n): every team size from 1 to 12 has a
    kernel (one thread per env up to 5 agents, G lanes per env above), checked against the float64 oracle over an
    episode from Philox resets with crowded starts; batch sizes that do not fill the last warp."""
    B, T = 1000 + n, 25
    env = _mk('simple_spread', n, B, precision, seed=100 + n)
    obs0 = env.reset()
    pos, vel, lm, _ = env.get_state()
    pos = pos * 0.35  # crowd the agents into a third of the arena so that contacts happen
    env.set_state(pos, vel, lm)
    spec = mpe_vec.Spec('simple_spread', n)
    v = mpe_vec.VecEnv(spec, B)
    v.set_state(_np(pos), _np(vel), _np(lm))
    assert obs0.shape == (B, n, 4 + 2 * n)
    _check(_np(env.observe()), v.observe(), precision, 'obs0', 1e-6)
    rng = np.random.RandomState(n)
    worst, contacts = np.zeros(B), 0
    for t in range(T):
        act = rng.randint(0, 5, (B, n)).astype(np.int32)
        obs, rew, done, info = env.step(act, info=True)
        o, r, (coll, occ) = v.step(act)
        contacts += int((coll > 1).sum())
        ii = info['info_i'].cpu().numpy()
        if precision == 'fp64':
            assert np.abs(_np(obs) - o).max() <= F64_ATOL and np.abs(_np(rew) - r).max() <= F64_ATOL, t
            assert np.array_equal(ii[:, :n], coll) and np.array_equal(ii[:, n], occ)
        else:
            worst = np.maximum(worst, np.abs(_np(obs) - o)[:, :, 2:].max(axis=(1, 2)))
    if precision == 'fp32':  # stiff contacts amplify fp32 rounding for as long as they last (see test_gpu_actor.py)
        assert worst.max() <= 5e-4 and np.quantile(worst, 0.99) <= 5e-5 and np.median(worst) <= 2e-6, \
            (worst.max(), np.quantile(worst, 0.99), np.median(worst))
    assert n == 1 or contacts > 100
