"""CPU tests that EXECUTE the reference's own code for the path (not a restatement of it).

``python -m oracle.build_ref`` byte-compiles the reference's modules from /root/reference into oracle/_ref;
tests/_stubs/multiagent supplies the third-party package they import.  What runs here:

  * experiments/scenarios.py:124-192 ``make_env`` and :6-63 ``local_obs_*`` - the reference's functions, executed,
    against oracle/mpe_ref.py's restatement of them and against the committed golden trajectories, BIT for bit;
  * rls/replay_buffer.py:9-91 ``ReplayBuffer`` against oracle/replay_ref.py;
  * rls/model/ac_network*_gumbel.py ``ActorNetwork`` against the committed actor fixtures and oracle/actor_ref.py.
"""
import os

import numpy as np
import pytest

from oracle import actor_ref, build_ref, mpe_ref, replay_ref

pytestmark = pytest.mark.skipif(not build_ref.available(),
                                reason='oracle/_ref not built (python -m oracle.build_ref needs /root/reference)')

MPE_FILES = [('simple_spread', None), ('simple_spread', 6), ('simple_spread', 9), ('simple_spread', 12),
             ('simple_reference', None), ('simple_speaker_listener', None)]


@pytest.fixture(scope='module')
def ref_scenarios():
    from tests import _refloop
    S, _, _ = _refloop.use_reference('oracle')
    assert S.__file__.startswith(build_ref.OUT)
    yield S
    _refloop.purge()


def _gold(golden_dir, scenario, n):
    return np.load(os.path.join(golden_dir, 'mpe_%s%s.npz' % (scenario, '' if n is None else '_n%d' % n)))


def _actions(g, t, b, N, multi, dim_c):
    out = []
    for i in range(N):
        a = np.zeros(5); a[g['act_u'][t, b, i]] = 1.0
        if multi:
            c = np.zeros(dim_c); c[g['act_c'][t, b, i]] = 1.0
            a = np.concatenate([a, c])
        out.append(a)
    return out


@pytest.mark.parametrize('scenario,n', MPE_FILES)
def test_reference_make_env_equals_the_oracle_bit_for_bit(ref_scenarios, golden_dir, scenario, n):
    """The reference's make_env (its flags, its observation monkey-patch) on the golden states and actions gives
    exactly the observations / rewards / benchmark flags the oracle's make_env gave when the fixtures were made."""
    g = _gold(golden_dir, scenario, n)
    env = ref_scenarios.make_env(scenario, n=n, benchmark=False, discrete_action=True, local_observation=True)
    mine = mpe_ref.make_env(scenario, n=n)
    # the observation callback IS the reference's function (compiled from experiments/scenarios.py)
    code = env.observation_callback.__func__.__code__
    assert code.co_filename == 'reference/experiments/scenarios.py' and code.co_name.startswith('local_obs_')
    # make_env's construction flags (scenarios.py:171,191) and the surface main.py:51-58 reads
    assert env.shared_reward is False and env.force_discrete_action is True and env.discrete_action_space is True
    assert env.n == mine.n == g['pos0'].shape[1]
    assert [s.shape for s in env.observation_space] == [s.shape for s in mine.observation_space]
    multi = hasattr(env.action_space[0], 'high')
    if multi:
        assert (env.action_space[0].high + 1).tolist() == (mine.action_space[0].high + 1).tolist() == [5, 10]
    else:
        assert env.action_space[0].n == mine.action_space[0].n == 5
    bench = ref_scenarios.make_env(scenario, n=n, benchmark=True) if scenario == 'simple_spread' else None
    B = min(g['pos0'].shape[0], 16)
    T = g['act_u'].shape[0]
    for b in range(B):
        for e in (env, mine):
            mpe_ref.set_state(e, g['pos0'][b], g['vel0'][b], g['lm0'][b], g['goal0'][b])
        assert np.array_equal(np.stack(mpe_ref.get_obs(env)), g['obs0'][b])
        for t in range(T):
            act = _actions(g, t, b, env.n, multi, env.world.dim_c)
            o, r, d, info = env.step([a.copy() for a in act])
            o2, r2, d2, info2 = mine.step([a.copy() for a in act])
            assert np.array_equal(np.stack(o), g['obs'][t, b]) and np.array_equal(np.stack(o), np.stack(o2)), (b, t)
            assert np.array_equal(np.array(r), g['rew'][t, b]) and np.array_equal(np.array(r), np.array(r2)), (b, t)
            assert d == d2 == [False] * env.n and info == info2 == {'n': [{}] * env.n}
            if bench is not None and t % 6 == 0:
                pos = np.stack([a.state.p_pos for a in env.world.agents])
                vel = np.stack([a.state.p_vel for a in env.world.agents])
                mpe_ref.set_state(bench, pos, vel, g['lm0'][b])
                tup = [bench._get_info(a) for a in bench.world.agents]
                assert [x[1] for x in tup] == g['coll'][t, b].tolist() and tup[0][3] == g['occ'][t, b]
                assert [x[0] for x in tup] == g['rew'][t, b].tolist()


def test_reference_make_env_reset_uses_numpy_global_rng_in_upstream_order(ref_scenarios):
    """env.seed + env.reset through the reference's make_env == the oracle's, draw for draw (main.py:41-49)."""
    for scenario, n in MPE_FILES:
        a = ref_scenarios.make_env(scenario, n=n)
        b = mpe_ref.make_env(scenario, n=n)
        a.seed(12345678); oa = a.reset()
        b.seed(12345678); ob = b.reset()
        assert all(np.array_equal(x, y) for x, y in zip(oa, ob))


def test_reference_unsupported_scenario_prints_but_does_not_patch(ref_scenarios, capsys):
    """scenarios.py:151-164: multi_speaker_listener falls through to scenarios.load, which has no such script here
    (the reference leaves its stock, ragged observation in place - 'origin' - so its own actor cannot run it;
    fullobs_collect_treasure is covered in tests/test_treasure_oracle.py)."""
    with pytest.raises(ImportError):
        ref_scenarios.make_env('multi_speaker_listener')


def test_reference_replay_buffer_equals_restatement():
    build_ref.add_to_path()
    from rls.replay_buffer import ReplayBuffer
    rng = np.random.RandomState(0)
    ref, mine = ReplayBuffer(size=7), replay_ref.ReplayRing(7)
    for t in range(19):
        tr = (rng.randn(3, 4), np.eye(5)[rng.randint(5, size=3)], np.float64(rng.randn()), rng.randn(3, 4),
              np.float64(t % 2))
        ref.add(*tr); mine.add(*tr)
        assert len(ref) == len(mine) and ref._next_idx == mine._next_idx
    idx = [0, 6, 3, 3, 1]
    for a, b in zip(ref.sample_index(idx), mine.sample_index(idx)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('tag,model', [('spread_n3', False), ('reference', False), ('model_n6', True)])
def test_reference_actor_network_reproduces_the_committed_fixtures(golden_dir, tag, model):
    """The fixtures tests/golden/actor_*.npz are outputs of the reference's ActorNetwork; re-run it here (fp32 CPU)
    from the stored state_dict: the outputs must come back exactly, and the float64 restatement within rounding."""
    import torch
    build_ref.add_to_path()
    if model:
        from rls.model.ac_network_model_multi_gumbel import ActorNetwork
    else:
        from rls.model.ac_network_multi_gumbel import ActorNetwork
    g = np.load(os.path.join(golden_dir, 'actor_%s.npz' % tag))
    sd = {k[3:]: g[k] for k in g.files if k.startswith('sd/')}
    D = sd['dense1.module.weight'].shape[1]
    A = [sd['dense2_1.module.weight'].shape[0], sd['dense2_2.module.weight'].shape[0]] \
        if 'dense2_1.module.weight' in sd else sd['dense2.module.weight'].shape[0]
    net = ActorNetwork(input_dim=D, out_dim=A)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    with torch.no_grad():
        out = net.forward(torch.from_numpy(np.array(g['obs'], dtype='float32')))
    nxt = None
    if model:
        out, nxt = out
    heads = out if isinstance(out, list) else [out]
    want = actor_ref.forward(sd, g['obs'])
    for hi, h in enumerate(heads):
        assert np.abs(h.numpy() - g['logits%d' % hi]).max() <= 1e-6  # same code, maybe another BLAS blocking
        assert np.abs(h.numpy() - want['logits'][hi]).max() < 2e-6
    if model:
        assert np.abs(nxt.numpy() - want['next_state']).max() < 2e-6


def test_reference_run_loop_executes_on_the_stub_env(tmp_path, monkeypatch):
    """experiments/run.py:11-103 (the caller of the hot path) runs unchanged on the oracle-backed stub with an
    acting-only policy; every transition its ReplayBuffer stored replays exactly through the oracle.  The same
    harness drives the CUDA env in tests/test_gpu_reference_loop.py."""
    from tests import _refloop
    scen, run_mod, arglist = _refloop.use_reference('oracle')
    saved = arglist.num_episodes
    monkeypatch.chdir(tmp_path)
    os.makedirs('Models')
    keep = []

    class RandomActor(object):
        def __init__(self, actor, critic, memory, action_type='Discrete'):
            self.memory = memory
            keep.append(self)

        def get_exploration_action(self, obs_n):
            return np.eye(5, dtype=np.float32)[np.random.randint(5, size=(1, len(obs_n)))]

        def save_models(self, fname):
            open('./Models/' + fname + '_actor.pt', 'w').close()

    try:
        env = scen.make_env('simple_spread', n=6)
        env.seed(3)
        arglist.num_episodes = 2
        run_mod.run(env, None, None, RandomActor, 'simple_spread_n_agent_6_', 'Discrete', cnt=0)
    finally:
        arglist.num_episodes = saved
        _refloop.purge()
    assert _refloop.verify_memory(keep[0].memory, 'simple_spread', 6, obs_tol=1e-12, rew_tol=1e-12) == 50
    assert os.path.exists('Models/history_simple_spread_n_agent_6__0.pkl')
