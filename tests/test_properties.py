"""Property tests (tier T6): invariants of the physics that do not need the oracle.
CPU part runs on the oracle with hypothesis; the GPU part checks the kernels at ragged sizes."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import mpe_vec


def _rand_state(seed, B, N):
    rng = np.random.RandomState(seed)
    pos = rng.uniform(-1, 1, (B, N, 2))
    pos[:, 1:] = pos[:, :1] + rng.uniform(-0.4, 0.4, (B, N - 1, 2))  # crowded: contacts are common
    return pos, rng.uniform(-1, 1, (B, N, 2)), rng.uniform(-1, 1, (B, N, 2)), rng.randint(0, 5, (B, N))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.sampled_from([2, 3, 6]))
def test_oracle_momentum_symmetry_and_permutation_equivariance(seed, n):
    B = 8
    pos, vel, lm, act = _rand_state(seed, B, n)
    spec = mpe_vec.Spec('simple_spread', n)
    e = mpe_vec.VecEnv(spec, B); e.set_state(pos, vel, lm)
    obs, rew, _ = e.step(act)
    # f_a = -f_b: total momentum changes only by damping and the actions
    table = np.array([[0, 0], [5, 0], [-5, 0], [0, 5], [0, -5]], dtype=np.float64)
    assert np.allclose(e.vel.sum(1), 0.75 * vel.sum(1) + 0.1 * table[act].sum(1), atol=1e-9)
    # relabelling the agents permutes positions/obs/rewards the same way
    perm = np.random.RandomState(seed + 1).permutation(n)
    e2 = mpe_vec.VecEnv(spec, B); e2.set_state(pos[:, perm], vel[:, perm], lm)
    obs2, rew2, _ = e2.step(act[:, perm])
    assert np.allclose(e2.pos, e.pos[:, perm], atol=1e-12) and np.allclose(obs2, obs[:, perm], atol=1e-12)
    assert np.allclose(rew2, rew[:, perm], atol=1e-12)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10 ** 6), shift=st.sampled_from([0.25, -0.5, 2.0]))
def test_oracle_translation_invariance_of_relative_observations(seed, shift):
    B, n = 6, 3
    pos, vel, lm, act = _rand_state(seed, B, n)
    spec = mpe_vec.Spec('simple_spread', n)
    a = mpe_vec.VecEnv(spec, B); a.set_state(pos, vel, lm)
    b = mpe_vec.VecEnv(spec, B); b.set_state(pos + shift, vel, lm + shift)
    oa, ra, _ = a.step(act); ob, rb, _ = b.step(act)
    assert np.allclose(oa[:, :, 4:], ob[:, :, 4:], atol=1e-9) and np.allclose(oa[:, :, :2], ob[:, :, :2], atol=1e-9)
    assert np.allclose(ra, rb, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [3, 6, 9, 12])
def test_gpu_permutation_equivariance_fp64(n):
    import multiagent_rl_b200 as m
    B = 1003  # ragged
    pos, vel, lm, act = _rand_state(7, B, n)
    perm = np.random.RandomState(3).permutation(n)
    e1 = m.make_env('simple_spread', n=n, num_envs=B, batched=True, precision='fp64')
    e2 = m.make_env('simple_spread', n=n, num_envs=B, batched=True, precision='fp64')
    e1.set_state(pos, vel, lm); e2.set_state(pos[:, perm], vel[:, perm], lm)
    o1, r1, _, _ = e1.step(act); o2, r2, _, _ = e2.step(act[:, perm])
    # forces are summed in agent-index order, so a relabelling changes the rounding order only
    assert float((o2.cpu() - o1.cpu()[:, perm]).abs().max()) < 1e-10
    assert float((r2.cpu() - r1.cpu()[:, perm]).abs().max()) < 1e-10


@pytest.mark.gpu
def test_gpu_every_kernel_at_ragged_sizes():
    """tools/sanitize_target.py: every kernel, fp32 and fp64, sizes that are not multiples of any tile."""
    import runpy, os
    from tests.conftest import ROOT
    runpy.run_path(os.path.join(ROOT, 'tools', 'sanitize_target.py'), run_name='__main__')
