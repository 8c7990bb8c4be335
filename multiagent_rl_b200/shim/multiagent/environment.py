"""``multiagent.environment.MultiAgentEnv`` with upstream's constructor (the call at experiments/scenarios.py:179-190)
on top of ``BatchedMultiAgentEnv``: one env instance, the reference's list-of-numpy surface."""
from multiagent_rl_b200.env import BatchedMultiAgentEnv


class MultiAgentEnv(BatchedMultiAgentEnv):
    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None, post_step_callback=None, shared_viewer=True,
                 discrete_action=True, **batch_options):
        name = world.scenario_name
        if done_callback is not None or post_step_callback is not None:
            raise NotImplementedError('done/post_step callbacks are not part of the %s kernels' % name)
        obs_fn = getattr(observation_callback, '__func__', observation_callback)
        if getattr(obs_fn, '__name__', '') != 'local_obs_' + name:
            # make_env(local_observation=False) leaves the stock full observation in place (scenarios.py:151-164)
            raise NotImplementedError('only the reference\'s partial observation local_obs_%s is implemented; '
                                      'got %r' % (name, getattr(obs_fn, '__name__', obs_fn)))
        super(MultiAgentEnv, self).__init__(name, n=world.num_agents, benchmark=info_callback is not None,
                                            discrete_action=discrete_action, **batch_options)
        self.shared_reward = bool(world.collaborative)
        self.world.collaborative = self.shared_reward
