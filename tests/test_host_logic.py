"""CPU tests of the host-side logic: sharding, the stats reduction over a 2-rank gloo group,
the gym-space stand-ins and the torch mirror of the actor."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from multiagent_rl_b200 import distributed as D
from multiagent_rl_b200.networks import ActorNetwork
from multiagent_rl_b200.spaces import Box, Discrete, MultiDiscrete
from oracle import actor_ref


def test_shard_range_partitions_exactly():
    for total in (1, 7, 65536, 1048576 + 3):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (o0, n0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + n0 == o1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env(backend='gloo')
    off, n = D.shard_range(1000, r, w)
    rets = np.arange(off, off + n, dtype=np.float64) * 0.01 - 3.0  # one finished episode per env
    local = [rets.sum(), (rets ** 2).sum(), float(n), 25.0 * n]
    out = D.reduce_return_stats(local)
    tmax = D.max_over_ranks(1.0 + r)
    g = D.gather_replay(torch.full((2, 3), float(r)))
    q.put((r, out, tmax, g.shape[0], float(g.sum())))
    torch.distributed.destroy_process_group()


def test_return_stats_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allr = np.arange(1000) * 0.01 - 3.0
    for r, out, tmax, gn, gs in res:
        assert out['episodes'] == 1000 and out['steps'] == 25000
        assert np.isclose(out['mean_return'], allr.mean()) and np.isclose(out['std_return'], allr.std())
        assert tmax == 2.0 and gn == 4 and gs == 6.0


def test_reduce_without_process_group():
    out = D.reduce_return_stats([10.0, 60.0, 2.0, 50.0])
    assert out['mean_return'] == 5.0 and np.isclose(out['std_return'], np.sqrt(5.0))


def test_spaces_surface():
    assert Discrete(5).n == 5 and 0 <= Discrete(5).sample() < 5
    md = MultiDiscrete([[0, 4], [0, 9]])
    assert (md.high + 1).tolist() == [5, 10]  # main.py:52-54
    assert Box(-np.inf, np.inf, (10,)).shape[0] == 10


@pytest.mark.parametrize('D_,A,model', [(10, 5, False), (21, [5, 10], False), (16, 5, True)])
def test_torch_mirror_matches_oracle_and_reference_keys(D_, A, model):
    sd = actor_ref.init_state_dict(D_, A, 3, model_head=model)
    net = ActorNetwork(D_, A, model_head=model)
    assert sorted(net.state_dict().keys()) == sorted(sd.keys())
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    obs = np.random.RandomState(0).uniform(-1, 1, (7, 3, D_)).astype(np.float32)
    with torch.no_grad():
        out = net(torch.from_numpy(obs))
    want = actor_ref.forward(sd, obs)
    nxt = None
    if model:
        out, nxt = out
    heads = out if isinstance(out, list) else [out]
    for h, w in zip(heads, want['logits']):
        assert np.max(np.abs(h.numpy() - w)) < 2e-6
    if model:
        assert np.max(np.abs(nxt.numpy() - want['next_state'])) < 2e-6


def test_golden_actor_loads_into_torch_mirror(golden_dir):
    g = np.load(os.path.join(golden_dir, 'actor_spread_n3.npz'))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')}
    net = ActorNetwork(10, 5)
    net.load_state_dict(sd)  # the reference's own state_dict keys
    with torch.no_grad():
        out = net(torch.from_numpy(g['obs'].astype(np.float32)))
    assert np.max(np.abs(out.numpy() - g['logits0'])) < 1e-6


def test_episode_history_matches_the_reference_bookkeeping(tmp_path):
    """experiments/run.py:23-25,55-57,62-65,94-100 run by hand on one env == EpisodeHistory on a batch of envs,
    column by column; the pickle has the keys reward_plot*.py / reward_test_phase_csv.py read."""
    import pickle
    from multiagent_rl_b200.history import EpisodeHistory
    B, N, L, T = 3, 3, 5, 17
    rng = np.random.RandomState(0)
    rew = rng.normal(size=(T, B, N))
    # the reference loop, once per env
    want = []
    for b in range(B):
        episode_rewards, agent_rewards, step = [0.0], [[0.0] for _ in range(N)], 0
        for t in range(T):
            for i in range(N):
                episode_rewards[-1] += rew[t, b, i]
                agent_rewards[i][-1] += rew[t, b, i]
            step += 1
            if step >= L:
                step = 0
                episode_rewards.append(0)
                for a in agent_rewards:
                    a.append(0)
        want.append((episode_rewards, agent_rewards))
    h = EpisodeHistory(B, N, max_episode_len=L)
    h.add_rollout(torch.from_numpy(rew[:9]))
    for t in range(9, T):
        h.add_step(torch.from_numpy(rew[t]))
    hist = h.history(order='env')
    E = T // L
    for b in range(B):
        assert np.allclose(hist['reward_episodes'][b * E:(b + 1) * E], want[b][0][:E], atol=1e-12)
        for i in range(N):
            assert np.allclose(hist['reward_episodes_by_agents'][i][b * E:(b + 1) * E], want[b][1][i][:E], atol=1e-12)
    # trailing in-progress entry = env 0's partial episode, as the reference's lists end
    assert np.isclose(hist['reward_episodes'][-1], want[0][0][-1])
    path = h.save(str(tmp_path / 'history_simple_spread_0.pkl'))
    with open(path, 'rb') as fp:
        back = pickle.load(fp)
    assert set(back) == {'reward_episodes', 'reward_episodes_by_agents'} and len(back['reward_episodes_by_agents']) == N
    assert len(back['reward_episodes']) == B * E + 1


def test_acting_trainer_checkpoint_names_match_the_reference(tmp_path, monkeypatch):
    """rls/agent/multiagent/ddpg_gumbel_fix.py:221-236: './Models/' + fname + '_actor.pt', where the caller
    (experiments/run.py:117) has already put arglist.appx into fname - the prefix must not be applied twice."""
    from multiagent_rl_b200.actor import ActingTrainer
    monkeypatch.chdir(tmp_path)
    appx = 'scalability/madr/'  # rls/arglist.py:29
    os.makedirs(os.path.join('Models', appx))
    net = ActorNetwork(10, 5)
    tr = ActingTrainer(net, None, None, 'Discrete', device='cpu')
    fname = appx + 'simple_spread' + '_fin_' + str(0)
    tr.save_models(fname)
    assert os.path.exists('./Models/scalability/madr/simple_spread_fin_0_actor.pt')
    want = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        for p in net.parameters():
            p.zero_()
    tr.load_models(fname)
    for k, v in net.state_dict().items():
        assert torch.equal(v, want[k])


def test_reduce_counts_nonfinite_episodes_separately():
    out = D.reduce_return_stats([10.0, 60.0, 3.0, 75.0, 1.0])  # 3 episodes, one of them NaN (kept out of the sums)
    assert out['episodes'] == 3 and out['nonfinite_episodes'] == 1 and out['mean_return'] == 5.0
