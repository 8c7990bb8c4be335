"""``multiagent.environment`` - the oracle's MultiAgentEnv behind upstream's constructor signature, which is the
one experiments/scenarios.py:179-190 calls (reset/reward/observation/post_step/info callbacks, discrete_action)."""
from oracle import mpe_ref


class MultiAgentEnv(mpe_ref.MultiAgentEnv):
    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None, post_step_callback=None, shared_viewer=True,
                 discrete_action=True):
        # version ambiguity (3) of oracle/mpe_ref.py: main.py:57 builds ONE head of width action_space[0].n for every
        # agent, so a world with a non-movable speaker (simple_speaker_listener) gets uniform Discrete(5) spaces
        uniform = 5 if any(not a.movable for a in world.agents) else None
        super(MultiAgentEnv, self).__init__(world, reset_callback=reset_callback, reward_callback=reward_callback,
                                            observation_callback=observation_callback, info_callback=info_callback,
                                            done_callback=done_callback, post_step_callback=post_step_callback,
                                            shared_viewer=shared_viewer, discrete_action=discrete_action,
                                            uniform_action_width=uniform)

    def render(self, mode='human', close=False):
        return []
