"""GPU tests of the device replay ring against oracle/replay_ref.py (the reference ReplayBuffer's semantics):
ring overwrite order, shared-reward sum, one-hot actions, gather by index, uniform sampling."""
import numpy as np
import pytest
import torch

from oracle import replay_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('N,D,A', [(3, 10, [5]), (2, 21, [5, 10]), (12, 28, [5])])
def test_ring_matches_reference_semantics(N, D, A):
    import multiagent_rl_b200 as m
    cap, B = 1000, 333
    buf = m.DeviceReplayBuffer(cap, N, D, A, seed=1)
    ring = replay_ref.ReplayRing(cap)
    rng = np.random.RandomState(0)
    for step in range(5):  # 1665 transitions through a ring of 1000: wraps
        obs = rng.randn(B, N, D).astype(np.float32); nxt = rng.randn(B, N, D).astype(np.float32)
        rew = rng.randn(B, N).astype(np.float32); done = (rng.rand(B) < 0.1).astype(np.float32)
        au = rng.randint(0, A[0], (B, N)); ac = rng.randint(0, A[1], (B, N)) if len(A) > 1 else None
        onehot = np.eye(A[0], dtype=np.float32)[au]
        if ac is not None:
            onehot = np.concatenate([onehot, np.eye(A[1], dtype=np.float32)[ac]], axis=-1)
        buf.add(obs, au, rew, nxt, done, act_c=ac)
        replay_ref.add_batched_step(ring, obs, onehot, rew, nxt, done)
        assert len(buf) == len(ring) and buf._next_idx == ring._next_idx
    idx = rng.randint(0, cap, 257)
    got = buf.sample_index(idx)
    want = ring.sample_index(idx)
    assert np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[3].cpu().numpy(), want[3])
    assert np.array_equal(got[1].cpu().numpy(), want[1])
    assert np.allclose(got[2].cpu().numpy(), want[2], atol=1e-6)  # fp32 sum over agents in agent order
    assert np.array_equal(got[4].cpu().numpy(), want[4])
    allv = buf.collect()
    assert allv[0].shape == (cap, N, D)
    buf.clear()
    assert len(buf) == 0


def test_uniform_sampling_and_errors():
    import multiagent_rl_b200 as m
    buf = m.DeviceReplayBuffer(5000, 3, 10, 5, seed=7)
    with pytest.raises(RuntimeError, match='empty'):
        buf.sample(4)
    z = torch.zeros(3000, 3, 10, device='cuda')
    buf.add(z, torch.zeros(3000, 3, dtype=torch.int32), torch.arange(9000, dtype=torch.float32).reshape(3000, 3), z)
    i1, i2 = buf.make_index(200_000), buf.make_index(200_000)
    assert int(i1.min()) >= 0 and int(i1.max()) < 3000 and not torch.equal(i1, i2)
    hist = torch.bincount(i1, minlength=3000).float()
    assert abs(float(hist.mean()) - 200_000 / 3000) < 1e-3 and float(hist.std()) < 12  # ~sqrt(66.7) = 8.2
    obs, act, rew, nxt, done = buf.sample(1024)
    assert obs.shape == (1024, 3, 10) and act.shape == (1024, 3, 5) and bool((act.sum(-1) == 1).all())
    assert bool(((rew / 3 - (rew / 3).round()).abs() < 1e-3).all())  # rew_shared = 3*(3i+1): every i in range
    latest = buf.make_latest_index(10)
    assert sorted(latest.tolist()) == list(range(2990, 3000))


def test_rollout_feeds_replay():
    """env.rollout(record=True) + replay: the transition stream of experiments/run.py:36-65 without leaving HBM."""
    import multiagent_rl_b200 as m
    from oracle import actor_ref
    B, T = 2048, 6
    env = m.make_env('simple_spread', num_envs=B, batched=True, seed=5, max_episode_len=25)
    actor = m.FusedActor(actor_ref.init_state_dict(10, 5, 0), seed=5)
    buf = m.DeviceReplayBuffer(B * T, 3, 10, 5)
    obs = env.reset()
    nxt, rew, au, _ = env.rollout(actor, T, record=True)
    prev = obs
    for t in range(T):
        buf.add(prev, au[t], rew[t], nxt[t])
        prev = nxt[t]
    assert len(buf) == B * T
    o, a, r, n, d = buf.sample_index(torch.arange(B, 2 * B))  # the transitions of step 1
    assert torch.equal(o, nxt[0]) and torch.equal(n, nxt[1]) and torch.allclose(r, rew[1].sum(1), atol=1e-5)
    assert torch.equal(a.argmax(-1).int(), au[1]) and float(d.abs().max()) == 0.0


def test_replay_rejects_mismatched_tensors_and_bad_indices():
    import multiagent_rl_b200 as m
    rb = m.DeviceReplayBuffer(64, 3, 10, 5)
    obs = torch.zeros((8, 3, 10), device='cuda')
    act = torch.zeros((8, 3), dtype=torch.int32, device='cuda')
    rew = torch.zeros((8, 3), device='cuda')
    with pytest.raises(ValueError):
        rb.add(obs, act[:4], rew, obs)
    with pytest.raises(ValueError):
        rb.add(obs, act, rew[:, :2], obs)
    with pytest.raises(ValueError):
        rb.add(obs[:, :2], act, rew, obs)
    rb.add(obs + torch.arange(8, device='cuda').view(8, 1, 1), act, rew, obs)
    with pytest.raises(IndexError):
        rb.sample_index([0, 8])
    with pytest.raises(IndexError):
        rb.sample_index(np.array([-9]))
    o = rb.sample_index([-1, 0])[0]   # Python list indexing: -1 is the newest transition
    assert float(o[0, 0, 0]) == 7.0 and float(o[1, 0, 0]) == 0.0
    # device-resident indices cannot raise without a sync: out-of-range values are clamped into the ring
    o = rb.sample_index(torch.tensor([1000, -1000, -2], device='cuda'))[0]
    assert [float(x) for x in o[:, 0, 0]] == [7.0, 0.0, 6.0]
