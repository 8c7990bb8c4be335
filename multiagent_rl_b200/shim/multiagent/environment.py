"""``multiagent.environment.MultiAgentEnv`` with upstream's constructor (the call at experiments/scenarios.py:179-190)
on top of ``BatchedMultiAgentEnv``: one env instance, the reference's list-of-numpy surface."""
from multiagent_rl_b200.env import BatchedMultiAgentEnv


class MultiAgentEnv(BatchedMultiAgentEnv):
    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None, post_step_callback=None, shared_viewer=True,
                 discrete_action=True, **batch_options):
        name = world.scenario_name
        # fullobs_collect_treasure hands its own post_step hook over (experiments/scenarios.py:174-177): the kernel
        # runs it (pick-up, respawn, deposit); any other hook has no kernel
        own_hook = name == 'fullobs_collect_treasure' and getattr(post_step_callback, '__name__', '') == 'post_step'
        if done_callback is not None or (post_step_callback is not None and not own_hook):
            raise NotImplementedError('done/post_step callbacks are not part of the %s kernels' % name)
        obs_fn = getattr(observation_callback, '__func__', observation_callback)
        expected = {'fullobs_collect_treasure': 'local_obs_collect_treasure'}.get(name, 'local_obs_' + name)
        if getattr(obs_fn, '__name__', '') != expected:
            # make_env(local_observation=False) leaves the stock full observation in place (scenarios.py:151-164)
            raise NotImplementedError('only the reference\'s partial observation local_obs_%s is implemented; '
                                      'got %r' % (name, getattr(obs_fn, '__name__', obs_fn)))
        super(MultiAgentEnv, self).__init__(name, n=world.num_agents, benchmark=info_callback is not None,
                                            discrete_action=discrete_action, **batch_options)
        self.shared_reward = bool(world.collaborative)
        self.world.collaborative = self.shared_reward
