"""GPU tests that run the reference's OWN entry path - experiments/scenarios.py ``make_env``, experiments/run.py
``run`` / ``run_test``, rls/agent/multiagent/{ddpg,model_ddpg}_gumbel_fix.py ``Trainer`` - unchanged (compiled
into oracle/_ref from /root/reference) on top of the CUDA drop-in:

  * ``multiagent`` resolves to multiagent_rl_b200/shim, so the reference's make_env builds the CUDA env;
  * ``class Trainer(FusedActingMixin, <reference Trainer>)`` swaps get_exploration_action for the fused kernel
    (and the plain reference Trainer is run too: its torch actor on the CUDA env).

After the loop ends, every transition the reference's ReplayBuffer recorded is replayed through the float64
oracle (fp32 env tolerance: 5e-5 on observations, 2e-4 per agent on the summed reward).
"""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import build_ref
from tests import _refloop

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason='oracle/_ref not built (python -m oracle.build_ref)')]


class _Numpy1(object):
    """numpy >= 2 raises on ``np.array(a_list, copy=False)``; the reference's ReplayBuffer._encode_sample
    (rls/replay_buffer.py:44-47) was written for numpy 1, where that meant "copy only if needed".  Environment
    compatibility for the reference's learner, nothing on the path under test."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def array(x, *a, copy=True, **kw):
        return np.asarray(x, *a, **kw) if copy is False else np.array(x, *a, copy=copy, **kw)


@pytest.fixture
def ref(tmp_path, monkeypatch):
    scen, run_mod, arglist = _refloop.use_reference('cuda')
    import rls.replay_buffer
    monkeypatch.setattr(rls.replay_buffer, 'np', _Numpy1())
    saved = {k: getattr(arglist, k) for k in dir(arglist) if not k.startswith('_')}
    monkeypatch.chdir(tmp_path)
    os.makedirs(os.path.join('Models', arglist.appx))
    arglist.actor_learning_rate = 1e-2   # main.py:34-35
    arglist.critic_learning_rate = 1e-2
    yield scen, run_mod, arglist
    for k, v in saved.items():
        setattr(arglist, k, v)
    _refloop.purge()


def _trainer_class(base, fused, keep):
    from multiagent_rl_b200 import FusedActingMixin
    bases = (FusedActingMixin, base) if fused else (base,)

    class Trainer(*bases):
        def __init__(self, *a, **kw):
            super(Trainer, self).__init__(*a, **kw)
            keep.append(self)
    return Trainer


@pytest.mark.parametrize('fused', [True, False])
def test_reference_run_trains_on_the_cuda_env(ref, fused):
    """experiments/run.py:11-103 ``run`` for 3 episodes with optimize() every 10 steps: the reference's loop, replay
    buffer, learner and checkpoint writer, the CUDA env underneath, acting through the fused kernel (or the stock
    torch actor)."""
    import multiagent_rl_b200 as m
    scen, run_mod, arglist = ref
    from rls.agent.multiagent.ddpg_gumbel_fix import Trainer as RefTrainer
    env = scen.make_env('simple_spread', benchmark=False, discrete_action=True, local_observation=True)
    assert isinstance(env, m.BatchedMultiAgentEnv) and env.n == 3 and env.shared_reward is False
    actor, critic, action_type = _refloop.main_py_setup(env, 12345678)
    arglist.num_episodes, arglist.warmup_steps, arglist.batch_size, arglist.update_rate = 3, 30, 16, 10
    arglist.save_rate = 2
    keep = []
    w0 = actor.dense1.module.weight.detach().clone()
    run_mod.run(env, actor, critic, _trainer_class(RefTrainer, fused, keep), 'simple_spread', action_type, cnt=0)
    learner = keep[0]
    # 3 full episodes of 25 steps, stored by the reference's ReplayBuffer
    assert len(learner.memory) == 75
    assert _refloop.verify_memory(learner.memory, 'simple_spread', None) == 75
    # optimize() ran (steps 40, 50, 60, 70) and the fused actor followed the weights
    assert not torch.equal(w0.to(learner.actor.dense1.module.weight.device), learner.actor.dense1.module.weight)
    if fused:
        assert learner._fused is not None and learner._fused_version == learner._weights_version()
        assert learner._act_step == 75
    # what run() leaves behind (run.py:94-102)
    with open('Models/history_simple_spread_0.pkl', 'rb') as fp:
        hist = pickle.load(fp)
    assert len(hist['reward_episodes']) == 4 and len(hist['reward_episodes_by_agents']) == 3
    rews = np.array([tr[2] for tr in learner.memory._storage]).reshape(3, 25).sum(1)
    assert np.allclose(hist['reward_episodes'][:3], rews, atol=1e-9)
    assert os.path.exists('Models/simple_spread_fin_0_actor.pt') and os.path.exists('Models/simple_spread_fin_0_critic.pt')


def test_reference_run_test_loads_a_checkpoint_and_rolls_out(ref):
    """main.py:63-65 with TEST_ONLY: run_test -> Trainer.load_models(arglist.appx + name) -> 20 episodes (500 steps,
    19 env.reset() calls in between), every transition replayed through the float64 oracle."""
    scen, run_mod, arglist = ref
    from rls.agent.multiagent.ddpg_gumbel_fix import Trainer as RefTrainer
    env = scen.make_env('simple_spread', benchmark=False, discrete_action=True, local_observation=True)
    actor, critic, action_type = _refloop.main_py_setup(env, 12345679)
    keep = []
    T = _trainer_class(RefTrainer, True, keep)
    T(actor, critic, None, action_type).save_models(arglist.appx + 'simple_spread_fin_0')
    saved = {k: v.clone() for k, v in actor.state_dict().items()}
    with torch.no_grad():
        for p in actor.parameters():
            p.add_(1.0)  # load_models must bring the saved weights back
    arglist.num_episodes = 20
    keep.clear()
    run_mod.run_test(env, actor, critic, T, 'simple_spread', action_type, cnt=0)
    learner = keep[0]
    for k, v in learner.actor.state_dict().items():
        assert torch.equal(v.cpu(), saved[k].cpu()), k
    assert len(learner.memory) == 500 and _refloop.verify_memory(learner.memory, 'simple_spread', None) == 500
    with open('Models/test_history_simple_spread_0.pkl', 'rb') as fp:
        hist = pickle.load(fp)
    assert len(hist['reward_episodes']) == 21 and len(hist['memory']) == 500
    # every episode starts from a fresh numpy-drawn state: the first observations of consecutive episodes differ
    starts = [np.stack(learner.memory._storage[25 * e][0]) for e in range(20)]
    assert all(not np.array_equal(starts[e], starts[e + 1]) for e in range(19))


@pytest.mark.parametrize('n', [6, 12])
def test_reference_scalability_path_model_trainer(ref, n):
    """main_scalability_1.py:30-65: make_env(n=n_agent), the '+model' actor and model_ddpg_gumbel_fix.Trainer."""
    scen, run_mod, arglist = ref
    from rls.agent.multiagent.model_ddpg_gumbel_fix import Trainer as RefTrainer
    env = scen.make_env('simple_spread', n=n, benchmark=False, discrete_action=True, local_observation=True)
    assert env.n == n and env.observation_space[0].shape[0] == 4 + 2 * n
    actor, critic, action_type = _refloop.main_py_setup(env, 12345678, model=True)
    arglist.num_episodes, arglist.warmup_steps, arglist.batch_size, arglist.update_rate = 2, 30, 16, 20
    keep = []
    name = 'simple_spread_n_agent_%d_' % n
    run_mod.run(env, actor, critic, _trainer_class(RefTrainer, True, keep), name, action_type, cnt=0)
    learner = keep[0]
    assert len(learner.memory) == 50 and _refloop.verify_memory(learner.memory, 'simple_spread', n) == 50
    assert learner._fused.has_model and os.path.exists('Models/%s_fin_0_actor.pt' % name)


@pytest.mark.parametrize('scenario', ['simple_reference', 'simple_speaker_listener'])
def test_reference_run_on_the_communication_scenarios(ref, scenario):
    """main.py:24-25,51-58: MultiDiscrete heads (simple_reference) and the width-5 uniform head of the
    speaker/listener pair, through run.py:36-44."""
    scen, run_mod, arglist = ref
    from rls.agent.multiagent.ddpg_gumbel_fix import Trainer as RefTrainer
    env = scen.make_env(scenario, benchmark=False, discrete_action=True, local_observation=True)
    actor, critic, action_type = _refloop.main_py_setup(env, 12345678)
    assert action_type == ('MultiDiscrete' if scenario == 'simple_reference' else 'Discrete')
    arglist.num_episodes = 2
    keep = []
    run_mod.run(env, actor, critic, _trainer_class(RefTrainer, True, keep), scenario, action_type, cnt=0)
    mem = keep[0].memory
    assert len(mem) == 50
    width = 15 if scenario == 'simple_reference' else 5
    for obs_n, action_n, rew, new_obs_n, done in mem._storage:
        assert len(action_n) == 2 and all(a.shape == (width,) for a in action_n)
        assert np.isfinite(rew) and rew <= 0.0 and done == 0.0
        o, o2 = np.stack(obs_n), np.stack(new_obs_n)
        # landmarks are static: the listener's relative landmark positions move by exactly -(its displacement)
        i = 1
        d = (o2[i, 2:8] - o[i, 2:8]).reshape(3, 2)
        assert np.abs(d - d[0]).max() < 1e-5
        if scenario == 'simple_reference':  # the message agent 0 sent is what agent 1 observes next
            assert np.array_equal(o2[1, 11:21], action_n[0][5:15]) and np.array_equal(o2[0, 11:21], action_n[1][5:15])


def test_reference_bicnet_baseline_loop(ref):
    """main.py:15-18: the BiCNet baseline (experiments/run_BIC.py, BIC_gumbel_fix.Trainer, per-agent rewards and dones
    in the replay tuples) - same actor architecture, so the fused acting path serves it too."""
    scen, _, arglist = ref
    import experiments.run_BIC as run_bic
    from rls.agent.multiagent.BIC_gumbel_fix import Trainer as RefTrainer
    env = scen.make_env('simple_spread', benchmark=False, discrete_action=True, local_observation=True)
    actor, critic, action_type = _refloop.main_py_setup(env, 12345678, bic=True)
    arglist.num_episodes = 2
    keep = []
    run_bic.run(env, actor, critic, _trainer_class(RefTrainer, True, keep), 'simple_spread', action_type, cnt=0)
    mem = keep[0].memory
    assert len(mem) == 50 and _refloop.verify_memory(mem, 'simple_spread', None) == 50
    assert len(mem._storage[0][2]) == 3 and mem._storage[0][4] == [0.0, 0.0, 0.0]


def test_reference_optimize_learns_from_the_device_replay(ref):
    """examples/train_batched.py: batched fused rollouts -> DeviceReplayBuffer -> the reference's Trainer.optimize()
    (ddpg_gumbel_fix.py:131-219, unchanged) -> weights back into the fused actor.  The learner's `memory` is the device
    ring, so process_batch (make_index / sample_index) runs on the GPU replay."""
    import importlib.util
    from tests.conftest import ROOT
    spec = importlib.util.spec_from_file_location('train_batched', os.path.join(ROOT, 'examples', 'train_batched.py'))
    tb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tb)
    lines = []
    returns, learner, hist = tb.train(num_envs=512, episodes=4, updates_per_episode=3, batch_size=256, log=lines.append)
    assert len(returns) == 4 and all(np.isfinite(returns)) and len(lines) == 4
    assert len(learner.memory) == 4 * 25 * 512
    h = hist.history()
    assert len(h['reward_episodes']) == 4 * 512 + 1
    # the replay's transitions chain: obs_next of step t is obs of step t + 1 inside an episode
    s0, a0, r, s1, d = learner.memory.sample_index(torch.arange(0, 1024, device='cuda'))
    assert torch.equal(s1[:512], s0[512:1024]) and bool((a0.sum(-1) == 1).all()) and float(d.abs().max()) == 0.0
    # the critic the learner trained evaluates on the device kernel too
    import multiagent_rl_b200 as m
    q_ref = learner.critic.forward(s0.to(learner.device), a0.to(learner.device)).detach()
    q = m.FusedCritic(learner.critic.state_dict(), obs_dim=10).forward(s0, a0)
    assert float((q - q_ref).abs().max()) <= 1e-4 * max(1.0, float(q_ref.abs().max()))
