"""GPU parity of fullobs_collect_treasure (csrc/env_treasure.cuh) against oracle/maac_ref.py and the
committed fixture tests/golden/mpe_fullobs_collect_treasure.npz, through the C ABI.

Tolerances.  fp64 build: 1e-12 abs on positions / observations / rewards, the state word (types, alive, holding) and
the benchmark flags bit-exact.  fp32 build, one step from the fixture's float64 state: 2e-5 on positions and
observations, 2e-4 on velocities (the crafted states put collectors millimetres apart: rounding the state to fp32
turns the contact normal by 6e-8 / dist, times a force of ~10, times dt), 2e-5 on rewards; random-action runs: 3e-6.
A state word / flag / treasure order may differ only where the deciding float64 distance lies within 1e-5 (1e-6) of
its threshold (or of the next treasure's distance)."""
import os

import numpy as np
import pytest
import torch

from oracle import actor_ref, build_ref, maac_ref, maac_vec, mpe_ref, philox

pytestmark = pytest.mark.gpu
SC = 'fullobs_collect_treasure'


def _gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'mpe_fullobs_collect_treasure.npz'))


def _make(B, precision='fp32', seed=1, **kw):
    import multiagent_rl_b200 as m
    return m.make_env(SC, num_envs=B, batched=True, precision=precision, seed=seed, **kw)


def _set(env, pos, vel, tr, flags):
    goal = np.zeros((len(flags), 8), dtype=np.int32)
    goal[:, 0] = flags
    env.set_state(pos, vel, tr, goal)


def _flags(env):
    return env.get_state()[3][:, 0].cpu().numpy().astype(np.int64)


def test_dims_and_spaces():
    env = _make(5)
    assert env.n == 8 and env.num_landmarks == 6 and env.obs_dim == 30 and env.act_c == 0
    assert all(s.n == 5 for s in env.action_space) and env.observation_space[0].shape == (30,)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_philox_reset_is_bit_exact(precision):
    B, seed = 1000, 77
    env = _make(B, precision, seed=seed, env_id_offset=10_000)
    gid = np.arange(B) + 10_000
    for episode in range(2):
        obs = env.reset()
        pos, vel, tr, goal = env.get_state()
        a, t, ty = philox.treasure_reset(seed, gid, episode)
        assert np.array_equal(pos.cpu().numpy(), a) and not vel.any()
        if precision == 'fp64':
            assert np.array_equal(tr.cpu().numpy(), t)
        else:
            assert np.array_equal(tr.cpu().numpy(), (t / 0.95).astype(np.float32) * np.float32(0.95))
        want = np.array([maac_ref.pack_flags(ty[b], [True] * 6, [-1] * 6) for b in range(B)])
        assert np.array_equal(goal[:, 0].cpu().numpy(), want)
        # the observation of a fresh episode: own position, zero velocity, nothing held
        o = obs.cpu().numpy()
        assert np.array_equal(o[:, :, 0:2], pos.cpu().numpy()) and not o[:, :, 2:6].any()


def test_fp64_build_reproduces_the_fixture(golden_dir):
    """25 steps from the crafted states (pick-ups, deposits, respawns from the Philox stream, the three mass
    pairings in contact, max_speed clipping): every channel against the loop oracle's committed output."""
    g = _gold(golden_dir)
    T, B = g['act_u'].shape[:2]
    env = _make(B, 'fp64', seed=int(g['seed']))
    env.reset()                      # episode 0, step counter 0: the keys of the fixture's respawn draws
    _set(env, g['pos0'], g['vel0'], g['tr0'], g['flags0'])
    assert np.abs(env.observe().cpu().numpy() - g['obs0']).max() <= 1e-12
    for t in range(T):
        obs, rew, done, info = env.step(torch.from_numpy(g['act_u'][t].astype(np.int32)), info=True)
        pos, vel, tr, goal = env.get_state()
        assert np.abs(pos.cpu().numpy() - g['pos'][t]).max() <= 1e-12, t
        assert np.abs(vel.cpu().numpy() - g['vel'][t]).max() <= 1e-12, t
        assert np.abs(tr.cpu().numpy() - g['tr'][t]).max() <= 1e-12, t
        assert np.array_equal(goal[:, 0].cpu().numpy(), g['flags'][t]), t
        assert np.abs(obs.cpu().numpy() - g['obs'][t]).max() <= 1e-12, t
        assert np.abs(rew.cpu().numpy() - g['rew'][t]).max() <= 1e-12, t
        assert np.array_equal(info['info_i'][:, :8].cpu().numpy(), g['info'][t]) and not done.any()


def _near_threshold(pos, tr, eps):
    """per env: does any collector-treasure / collector-collector / collector-deposit distance lie within eps of
    its contact threshold, or do two treasure distances of one agent lie within eps of each other?"""
    B = pos.shape[0]
    near = np.zeros(B, dtype=bool)
    eps = np.broadcast_to(np.asarray(eps, dtype=np.float64), (B,))[:, None, None]   # scalar or one band per env
    c, d = pos[:, :6], pos[:, 6:]
    dct = np.linalg.norm(c[:, :, None] - tr[:, None], axis=-1)
    near |= (np.abs(dct - 0.075) < eps).any((1, 2))
    dcc = np.linalg.norm(c[:, :, None] - c[:, None], axis=-1)
    near |= (np.abs(dcc - 0.1) < eps).any((1, 2))
    dcd = np.linalg.norm(c[:, :, None] - d[:, None], axis=-1)
    near |= (np.abs(dcd - 0.125) < eps).any((1, 2))
    dat = np.sort(np.linalg.norm(pos[:, :, None] - tr[:, None], axis=-1), axis=-1)
    near |= (np.diff(dat, axis=-1) < eps).any((1, 2))
    return near


def test_fp32_build_one_step_from_every_fixture_state(golden_dir):
    g = _gold(golden_dir)
    T, B = g['act_u'].shape[:2]
    env = _make(B, 'fp32', seed=int(g['seed']))
    env.reset()
    prev = (g['pos0'], g['vel0'], g['tr0'], g['flags0'])
    checked = 0
    for t in range(T):
        _set(env, *prev)             # the fixture's float64 state, rounded to fp32 (step counter keeps running)
        obs, rew, done, info = env.step(torch.from_numpy(g['act_u'][t].astype(np.int32)), info=True)
        pos, vel, tr, goal = env.get_state()
        ok = ~_near_threshold(g['pos'][t], prev[2], 1e-5)
        # respawned treasures: positions are drawn in fp32 (u * 0.95f)
        assert np.abs(pos.cpu().numpy() - g['pos'][t])[ok].max() <= 2e-5, t
        assert np.abs(vel.cpu().numpy() - g['vel'][t])[ok].max() <= 2e-4, t
        assert np.abs(tr.cpu().numpy() - g['tr'][t])[ok].max() <= 1e-4, t      # -999 and fresh draws
        assert np.array_equal(goal[:, 0].cpu().numpy()[ok], g['flags'][t][ok]), t
        o = obs.cpu().numpy()
        far = np.abs(g['obs'][t]) > 100                                          # offsets to a dead treasure (~ -999)
        assert np.abs(o - g['obs'][t])[ok][~far[ok]].max() <= 2e-4, t   # rows carry the velocities
        assert not far[ok].any() or np.abs(o - g['obs'][t])[ok][far[ok]].max() <= 1e-4, t
        assert np.abs(rew.cpu().numpy() - g['rew'][t])[ok].max() <= 2e-5 + 1e-4 * far[ok].any(), t
        assert np.array_equal(info['info_i'][:, :8].cpu().numpy()[ok], g['info'][t][ok])
        checked += int(ok.sum())
        prev = (g['pos'][t], g['vel'][t], g['tr'][t], g['flags'][t])
    assert checked >= 0.95 * T * B


@pytest.mark.parametrize('precision,B', [('fp64', 96), ('fp32', 129)])
def test_long_run_against_the_loop_oracle(precision, B):
    """Three 25-step episodes from the kernels' own Philox resets with random actions, ragged batch (partial warp):
    the loop oracle is stepped env by env with the same draws.  fp32 is re-synchronised from the device state before
    every step (single-step comparison); fp64 runs free."""
    seed = 4242
    env = _make(B, precision, seed=seed, env_id_offset=500)
    rng = np.random.RandomState(5)
    oracles = []
    for b in range(B):
        e = mpe_ref.make_env(SC)
        e.scenario.draws = maac_ref.PhiloxDraws(seed, 500 + b)
        oracles.append(e)
    tol = 1e-12 if precision == 'fp64' else 3e-6
    events = 0
    for episode in range(3):
        obs = env.reset().cpu().numpy()
        for b, e in enumerate(oracles):
            e.scenario.draws.episode = episode
            o = e.reset()
            assert np.abs(np.stack(o) - obs[b]).max() <= (0 if precision == 'fp64' else 3e-7)
        for t in range(25):
            act = rng.randint(0, 5, (B, 8)).astype(np.int32)
            if precision == 'fp32':
                pos, vel, tr, goal = (x.cpu().numpy() for x in env.get_state())
            obs, rew, done, _ = env.step(torch.from_numpy(act))
            obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
            f_dev = _flags(env)
            for b, e in enumerate(oracles):
                e.scenario.draws.tstep = t
                if precision == 'fp32':
                    maac_ref.set_state(e, pos[b], vel[b], tr[b], goal[b, 0])
                f0 = maac_ref.get_flags(e)
                o, r, _, _ = e.step([np.eye(5)[k] for k in act[b]])
                events += int(maac_ref.get_flags(e) != f0)
                if precision == 'fp32' and _near_threshold(np.stack([a.state.p_pos for a in e.world.agents])[None],
                                                           tr[b][None], 1e-6)[0]:
                    continue
                far = np.abs(np.stack(o)) > 100
                assert np.abs(np.stack(o) - obs[b])[~far].max() <= tol, (episode, t, b)
                assert np.abs(np.array(r) - rew[b]).max() <= (tol if precision == 'fp64' else 2e-5 + 1e-4 * far.any())
                assert maac_ref.get_flags(e) == f_dev[b], (episode, t, b)
    assert events > 0


def test_masked_reset_and_list_surface_with_benchmark_info():
    """env.reset(mask) restarts only the masked envs (fresh Philox episode, nobody holds, every treasure alive) and
    leaves the others bit-untouched; the one-env list surface returns upstream's types (list of ndarray(30,), floats,
    [False] * 8, {'n': [0 / 1 per agent]} with benchmark=True) and turns the caller's action arrays into one-hots."""
    import multiagent_rl_b200 as m
    B = 203
    env = _make(B, seed=8)
    env.reset()
    g = torch.Generator().manual_seed(3)
    for t in range(6):
        env.step(torch.randint(0, 5, (B, 8), generator=g, dtype=torch.int32))
    before = [x.clone() for x in env.get_state()]
    mask = (torch.arange(B) % 3 == 0).to(torch.uint8)
    obs = env.reset(mask=mask)
    after = env.get_state()
    keep = ~mask.bool().cuda()
    for a, b in zip(after, before):
        assert torch.equal(a[keep], b[keep])
    a, t, ty = philox.treasure_reset(8, np.arange(B), 1)
    sel = mask.bool().numpy()
    assert np.array_equal(after[0].cpu().numpy()[sel], a[sel]) and not after[1][mask.bool().cuda()].any()
    want = np.array([maac_ref.pack_flags(ty[b], [True] * 6, [-1] * 6) for b in range(B)])
    assert np.array_equal(after[3][:, 0].cpu().numpy()[sel], want[sel])
    assert torch.equal(obs[:, :, 0:2], after[0])
    # list surface (num_envs = 1, numpy draws) with benchmark info
    one = m.make_env(SC, benchmark=True)
    one.seed(5)
    o = one.reset()
    assert len(o) == 8 and all(x.shape == (30,) and x.dtype == np.float64 for x in o)
    acts = [np.array([0.1, 0.7, 0.05, 0.1, 0.05]) for _ in range(8)]
    o, r, d, info = one.step(acts)
    assert all(a.tolist() == [0.0, 1.0, 0.0, 0.0, 0.0] for a in acts)          # in place, like upstream
    assert d == [False] * 8 and len(r) == 8 and all(isinstance(x, float) for x in map(float, r))
    assert set(info['n']) <= {0, 1} and len(info['n']) == 8
    assert (np.abs(np.stack(o)[:, 2] - 0.225) < 1e-6).sum() >= 6                 # accelerated to +x: 1.5 * 1.5 * 0.1 (unless in contact)


def test_sharding_is_invariant_to_the_number_of_shards():
    """envs shard in contiguous blocks, reset and respawn streams are keyed by the GLOBAL env id - one env
    of 96 instances and three shards of 32 with env_id_offset give bit-identical trajectories (30 steps: respawns occur)."""
    whole = _make(96, seed=21)
    parts = [_make(32, seed=21, env_id_offset=32 * k) for k in range(3)]
    o = whole.reset()
    assert torch.equal(o, torch.cat([p.reset() for p in parts]))
    g = torch.Generator().manual_seed(9)
    changed = 0
    for t in range(30):
        act = torch.randint(0, 5, (96, 8), generator=g, dtype=torch.int32)
        f0 = _flags(whole)
        o, r, _, _ = whole.step(act)
        outs = [p.step(act[32 * k:32 * k + 32]) for k, p in enumerate(parts)]
        assert torch.equal(o, torch.cat([x[0] for x in outs])) and torch.equal(r, torch.cat([x[1] for x in outs])), t
        assert np.array_equal(_flags(whole), np.concatenate([_flags(p) for p in parts]))
        changed += int((_flags(whole) != f0).sum())
    assert changed > 0


def test_tracked_returns_and_auto_reset():
    """rollout-style use: step + auto-reset after max_episode_len steps, episode statistics folded on the device."""
    B = 300
    env = _make(B, seed=9, max_episode_len=25)
    env.track_returns(True)
    obs = env.reset()
    tot = torch.zeros(B, dtype=torch.float64, device=obs.device)
    g = torch.Generator(device='cpu').manual_seed(0)
    for t in range(25):
        act = torch.randint(0, 5, (B, 8), generator=g, dtype=torch.int32)
        obs, rew, done, _ = env.step(act)
        tot += rew.double().sum(1)
    env.reset()
    st = env.read_stats()
    assert st[2] == B and st[3] == 25 * B and st[4] == 0
    assert abs(st[0] - float(tot.sum())) <= 1e-3 * B


def test_rollout_with_the_actor_kernels():
    """mpe_rollout for the 8-agent team: tensor-core actor (k_tc2<N = 8>) + step kernel per step; bit-identical to the
    stepwise calls, logits of the actor within 1e-5 of the float64 restatement, both actor paths agree."""
    import multiagent_rl_b200 as m
    B, T = 700, 30
    sd = actor_ref.init_state_dict(30, 5, 21)
    envs = [_make(B, seed=3, max_episode_len=25) for _ in range(2)]
    actor = m.FusedActor(sd, seed=3)
    obs = envs[0].reset()
    out = actor.forward(obs, want_logits=True)
    want = actor_ref.forward(sd, obs.cpu().numpy())['logits'][0]
    assert np.abs(out['logits'].cpu().numpy() - want).max() < 1e-5
    simt = m.FusedActor(sd, seed=3, impl='simt').forward(obs, want_logits=True)
    assert np.abs(simt['logits'].cpu().numpy() - want).max() < 1e-5
    rec = envs[0].rollout(actor, T, step0=0, record=True)
    obs = envs[1].reset()
    for t in range(T):
        a = actor.forward(obs, step=t)
        assert torch.equal(a['act_u'], rec[2][t]), t
        obs, rew, _, _ = envs[1].step(a['act_u'])
        assert torch.equal(obs, rec[0][t]) and torch.equal(rew, rec[1][t]), t
        if (t + 1) % 25 == 0:
            obs = envs[1].reset()
    assert torch.isfinite(rec[0]).all() and rec[1].abs().max() < 200


def test_rollout_at_65536_envs_vs_float64_oracle_directly():
    """BASELINE configs[1] size: a recorded 25-step mpe_rollout (tensor-core actor + step kernel) of 65,536 envs
    against the vectorised float64 oracle fed the recorded actions, from the kernel's own Philox reset.  Free-running
    fp32 vs float64 over an episode: positions 5e-4 worst case / 2e-5 at p99.9 (contacts between agents millimetres
    apart amplify rounding), rewards 2e-3; an env leaves the comparison once a contact flag or a neighbouring pair of
    the sorted treasure list lies within the band its own position error explains (3 x error + 1e-5)."""
    import multiagent_rl_b200 as m
    B, T, seed = 65_536, 25, 20240607
    env = _make(B, seed=seed, max_episode_len=25)
    actor = m.FusedActor(actor_ref.init_state_dict(30, 5, 4), seed=seed)
    obs0 = env.reset()
    v = maac_vec.VecTreasure(B, seed=seed)
    assert np.abs(v.reset() - obs0.cpu().numpy()).max() <= 3e-7
    rec = env.rollout(actor, T, step0=0, record=True)
    obs, rew, act = (x.cpu().numpy() for x in rec[:3])
    assert act.min() >= 0 and act.max() <= 4 and len(np.unique(act)) == 5
    clean = np.ones(B, dtype=bool)    # no contact flag near its threshold so far
    perr = []
    for t in range(T):
        tr_before = v.tr.copy()       # contacts and the sorted list see the treasures as they were BEFORE post_step
        for l in range(6):            # dead treasures all sit at (-999, -999): move them apart (exact ties go by index)
            tr_before[tr_before[:, l, 0] < -900, l] = 1000.0 * (l + 2)
        o, r, _ = v.step(act[t])
        # an env drops out once a deciding distance (contact thresholds, neighbouring entries of the sorted treasure
        # list) lies within the band its own fp32 position error explains
        err = np.abs(obs[t][:, :, 0:2] - v.pos).max((1, 2))
        clean &= ~_near_threshold(v.pos, tr_before, 3.0 * err + 1e-5)
        far = np.abs(o) > 100
        d = np.abs(o - obs[t])
        perr.append(d[clean][:, :, :4].max(-1).ravel())
        assert d[clean][~far[clean]].max() <= 5e-4, t
        assert np.abs(r - rew[t])[clean].max() <= 2e-3, t
    perr = np.concatenate(perr)
    assert np.quantile(perr, 0.999) <= 2e-5 and np.median(perr) <= 1e-6
    assert clean.mean() > 0.9
    f = env.get_state()[3][:, 0].cpu().numpy().astype(np.int64)
    # the kernel auto-reset after the 25th step (max_episode_len): compare the state word one step earlier instead
    assert (v.tstep == 25).all() and clean.sum() > 50_000 and f.shape == (B,)


def test_full_size_properties():
    """1,048,576 envs x 50 steps with random actions: size-independent invariants of the scenario -
    max_speed holds, positions of alive treasures stay inside the respawn box, a dead treasure sits at -999 for
    exactly one step, holders hold an existing type, the observation's treasure list is sorted, rewards are finite."""
    B = 1 << 20
    env = _make(B, seed=5)
    obs = env.reset()
    g = torch.Generator(device='cuda').manual_seed(1)
    dead_prev = torch.zeros(B, 6, dtype=torch.bool, device='cuda')
    picked = 0
    for t in range(50):
        act = torch.randint(0, 5, (B, 8), generator=g, device='cuda', dtype=torch.int32)
        obs, rew, done, _ = env.step(act)
        pos, vel, tr, goal = env.get_state()
        assert float(vel.norm(dim=-1).max()) <= 1.0 + 1e-6
        f = goal[:, 0]
        alive = ((f[:, None] >> (6 + torch.arange(6, device='cuda'))) & 1).bool()
        assert bool(((tr.abs().amax(-1) <= 0.95) == alive).all()) and bool((tr[~alive] == -999).all())
        assert not bool((dead_prev & ~alive).any())      # respawn_prob = 1: back after one step
        dead_prev = ~alive
        hold = (f[:, None] >> (12 + 2 * torch.arange(6, device='cuda'))) & 3
        assert int(hold.max()) <= 2 and int(f.max()) < (1 << 24)
        picked += int((~alive).sum())
        off = obs[:, :, 6:].reshape(B, 8, 6, 4)[..., :2]
        d2 = (off.double() ** 2).sum(-1)
        assert bool((d2[..., 1:] >= d2[..., :-1] * (1 - 1e-6)).all())
        assert bool(torch.isfinite(rew).all())
    assert picked > 1000


@pytest.mark.skipif(not build_ref.available(), reason='oracle/_ref not built (python -m oracle.build_ref)')
def test_reference_make_env_and_run_loop_on_the_cuda_env(tmp_path, monkeypatch):
    """The reference's own make_env('fullobs_collect_treasure') (experiments/scenarios.py:124-192) on the multiagent
    SHIM builds the CUDA env; its run loop (experiments/run.py:11-103) then drives it for two episodes with the fused
    acting mixin.  Every stored transition is replayed through the float64 oracle: state from the observation rows
    (own position / velocity / holding flags; treasure offsets of agent 0), one oracle step, same next observation
    and summed reward."""
    from tests import _refloop
    import multiagent_rl_b200 as m
    scen, run_mod, arglist = _refloop.use_reference('cuda')
    saved = {k: getattr(arglist, k) for k in dir(arglist) if not k.startswith('_')}
    monkeypatch.chdir(tmp_path)
    os.makedirs(os.path.join('Models', arglist.appx))
    keep = []
    try:
        from rls.agent.multiagent.ddpg_gumbel_fix import Trainer as RefTrainer

        class Trainer(m.FusedActingMixin, RefTrainer):
            def __init__(self, *a, **kw):
                super(Trainer, self).__init__(*a, **kw)
                keep.append(self)

        env = scen.make_env(SC, benchmark=False, discrete_action=True, local_observation=True)
        assert isinstance(env, m.BatchedMultiAgentEnv) and env.n == 8 and env.shared_reward is False
        actor, critic, action_type = _refloop.main_py_setup(env, 12345678)
        assert action_type == 'Discrete' and actor.dense1.module.weight.shape == (64, 30)
        arglist.num_episodes, arglist.warmup_steps, arglist.save_rate = 2, 10 ** 9, 1
        arglist.actor_learning_rate = arglist.critic_learning_rate = 1e-3   # main.py:29-31
        run_mod.run(env, actor, critic, Trainer, SC, action_type, cnt=0)
    finally:
        for k, v in saved.items():
            setattr(arglist, k, v)
        _refloop.purge()
    mem = keep[0].memory
    assert len(mem) == 50
    # Replay.  The observation a step returns is taken BEFORE post_step, so pick-ups / deposits / respawns are not in
    # it: the oracle runs in lock-step from the same numpy stream (main_py_setup seeded it; the env's reset and respawn
    # draws are its only consumers), keeps its own treasure / holding state, and takes the agents' positions and
    # velocities from the stored fp32 rows before every step.
    ora = mpe_ref.make_env(SC)
    np.random.seed(12345678)
    events = 0
    for k, (obs_n, action_n, rew_shared, new_obs_n, done) in enumerate(mem._storage):
        if k % 25 == 0:
            first = ora.reset()
            assert np.abs(np.stack(first) - np.stack(obs_n)).max() <= 3e-7, k   # same numpy draws, fp32 state
        o = np.stack(obs_n)
        w = ora.world
        for i, a in enumerate(w.agents):
            a.state.p_pos, a.state.p_vel = o[i, 0:2].astype(np.float64), o[i, 2:4].astype(np.float64)
        w.calculate_distances()
        f0 = maac_ref.get_flags(ora)
        o2, r2, _, _ = ora.step([np.array(a, dtype=np.float64) for a in action_n])
        events += int(maac_ref.get_flags(ora) != f0)
        far = np.abs(np.stack(o2)) > 100
        d = np.abs(np.stack(o2) - np.stack(new_obs_n))
        assert d[~far].max() <= 5e-6 and (not far.any() or d[far].max() <= 1e-4), k
        assert abs(np.sum(r2) - rew_shared) <= 1e-3, k
        assert all(a.shape == (5,) and a.sum() == 1.0 for a in action_n)
