"""Helpers for the tests that drive the reference's own loops (experiments/run.py) - TEST INFRASTRUCTURE.

``purge()`` / ``use_reference(...)`` control which ``multiagent`` package the compiled reference
(oracle/_ref) binds to: the CUDA shim (multiagent_rl_b200/shim) or the oracle-backed stub (tests/_stubs).
``verify_memory`` replays every transition the reference's ReplayBuffer recorded through the float64 oracle.
"""
import os
import sys

import numpy as np

from oracle import build_ref, mpe_ref
from tests.conftest import ROOT

STUBS = os.path.join(ROOT, 'tests', '_stubs')
_PREFIXES = ('multiagent', 'experiments')


def purge():
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + '.') for p in _PREFIXES):
            del sys.modules[k]
    for p in list(sys.path):
        if p == STUBS or p.endswith(os.path.join('multiagent_rl_b200', 'shim')):
            sys.path.remove(p)


def use_reference(backend):
    """backend: 'cuda' (the product shim) or 'oracle' (the test stub).  Returns the reference's modules."""
    purge()
    if backend == 'cuda':
        import multiagent_rl_b200.shim as shim
        shim.install()
    else:
        sys.path.insert(0, STUBS)
    build_ref.add_to_path()
    import experiments.run as run_mod
    import experiments.scenarios as scen_mod
    from rls import arglist
    return scen_mod, run_mod, arglist


def main_py_setup(env, seed, model=False, bic=False):
    """main.py:41-61 / main_scalability_1.py:38-58: seeds, dims, networks of the reference's own classes."""
    import torch
    if bic:  # main.py:15-18, the BiCNet baseline
        from rls.model.ac_network_multi_gumbel_BIC import ActorNetwork, CriticNetwork
    elif model:
        from rls.model.ac_network_model_multi_gumbel import ActorNetwork, CriticNetwork
    else:
        from rls.model.ac_network_multi_gumbel import ActorNetwork, CriticNetwork
    env.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    dim_obs = env.observation_space[0].shape[0]
    if hasattr(env.action_space[0], 'high'):
        dim_action = (env.action_space[0].high + 1).tolist()
        action_type = 'MultiDiscrete'
    else:
        dim_action = env.action_space[0].n
        action_type = 'Discrete'
    actor = ActorNetwork(input_dim=dim_obs, out_dim=dim_action)
    critic = CriticNetwork(input_dim=dim_obs + np.sum(dim_action), out_dim=1)
    return actor, critic, action_type


def state_from_obs(scenario, obs_n, goal=None):
    """Invert the reference's partial observations (experiments/scenarios.py:6-63) where they determine the state.
    simple_spread: obs = [vel, pos, landmarks - pos] -> everything.  The 2-agent scenarios' observations do not
    contain positions, so their transitions are checked through translation-invariant quantities instead."""
    assert scenario == 'simple_spread'
    o = np.stack(obs_n)
    vel, pos = o[:, 0:2], o[:, 2:4]
    lm = o[0, 4:].reshape(-1, 2) + pos[0]
    return pos, vel, lm


def verify_memory(memory, scenario, n, obs_tol=5e-5, rew_tol=2e-4):
    """Every (obs, action, sum(rew), obs', done) tuple in the reference's ReplayBuffer (rls/replay_buffer.py:30-37,
    filled at experiments/run.py:52) must be one float64 oracle step: state from obs, step with the stored one-hot
    action, compare obs' and the shared reward.  Returns the number of transitions checked."""
    ora = mpe_ref.make_env(scenario, n=n)
    count = 0
    for obs_n, action_n, rew_shared, new_obs_n, done in memory._storage:
        assert np.all(np.asarray(done) == 0.0)   # run.py stores float(done), run_BIC.py a list of floats
        for a in action_n:  # exact one-hots of width 5 (force_discrete_action's in-place rewrite)
            assert a.shape == (5,) and a.sum() == 1.0 and set(np.unique(a)) <= {0.0, 1.0}
        pos, vel, lm = state_from_obs(scenario, obs_n)
        mpe_ref.set_state(ora, pos, vel, lm)
        o, r, _, _ = ora.step([np.array(a, dtype=np.float64) for a in action_n])
        assert np.abs(np.stack(o) - np.stack(new_obs_n)).max() <= obs_tol, count
        assert abs(np.sum(r) - np.sum(rew_shared)) <= rew_tol * len(r), count  # run_BIC.py stores the per-agent list
        if np.ndim(rew_shared) == 1:
            assert np.abs(np.asarray(r) - np.asarray(rew_shared)).max() <= rew_tol, count
        count += 1
    return count
