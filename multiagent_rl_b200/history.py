"""Episode-return history of a batched run in the format the reference writes and its plotting scripts read.

``experiments/run.py:23-25,55-57,62-65,94-100`` keeps ``episode_rewards`` (sum of all agents' rewards per
episode) and ``agent_rewards[i]`` (per agent), appends a fresh ``0`` at every episode end and pickles

    {'reward_episodes': episode_rewards, 'reward_episodes_by_agents': agent_rewards}

to ``Models/history_<scenario>_<cnt>.pkl``; ``experiments/reward_plot*.py`` and ``reward_test_phase_csv.py:35-49``
read ``reward_episodes``.  ``EpisodeHistory`` builds the same object from the per-step reward tensors of the
batched environment (``step_tensor`` / ``rollout(record=True)``): the running sums live on the device in float64
(the reference accumulates Python floats), one [B, N] copy to the host per finished episode.  Plumbing, not a
kernel: plain torch ops on whatever device the rewards are on.
"""
import pickle

import numpy as np
import torch


class EpisodeHistory(object):
    def __init__(self, num_envs, n_agents, max_episode_len=25, device=None):
        self.B, self.N, self.L = int(num_envs), int(n_agents), int(max_episode_len)
        self.acc = torch.zeros((self.B, self.N), dtype=torch.float64, device=device)
        self.t = 0                # steps into the current episode (all envs step in lock-step)
        self.finished = []        # one float64 ndarray [B, N] per finished episode

    @classmethod
    def for_env(cls, env):
        return cls(env.num_envs, env.n, env.max_episode_len, device=env.device)

    def add_step(self, rew):
        """rew: [B, N] rewards of one step (experiments/run.py:55-57); closes the episode after max_episode_len
        steps like the ``terminal`` test at run.py:50,59."""
        self.acc += rew.to(self.acc.dtype).reshape(self.B, self.N)
        self.t += 1
        if self.t >= self.L:
            self.end_episode()

    def add_rollout(self, rew):
        """rew: [T, B, N] rewards recorded by ``env.rollout(actor, T, record=True)``."""
        for k in range(rew.shape[0]):
            self.add_step(rew[k])

    def end_episode(self):
        self.finished.append(self.acc.cpu().numpy().copy())
        self.acc.zero_()
        self.t = 0

    def history(self, order='episode'):
        """The reference's dict.  Batched episodes are flattened episode-major (all envs' episode 0, then episode 1,
        ...; ``order='env'``: env-major), followed by the in-progress entry the reference's lists always end with."""
        if self.finished:
            fin = np.stack(self.finished)                      # [E, B, N]
            if order == 'env':
                fin = fin.transpose(1, 0, 2)
            flat = fin.reshape(-1, self.N)
        else:
            flat = np.zeros((0, self.N))
        cur = self.acc.cpu().numpy()
        tail = cur.sum(0, keepdims=True) * 0.0 if self.t == 0 else cur[:1]  # a fresh 0, or env 0's partial episode
        flat = np.concatenate([flat, tail.reshape(1, self.N)], 0)
        return {'reward_episodes': [float(x) for x in flat.sum(1)],
                'reward_episodes_by_agents': [[float(x) for x in flat[:, i]] for i in range(self.N)]}

    def save(self, path, order='episode'):
        """``pickle.dump(hist, fp)`` (experiments/run.py:96-99)."""
        with open(path, 'wb') as fp:
            pickle.dump(self.history(order), fp)
        return path
