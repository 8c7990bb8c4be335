"""``multiagent.scenarios.load(name + '.py')`` (experiments/scenarios.py:150) -> an object with a ``Scenario`` class."""
from ..core import World

SUPPORTED = ('simple_spread', 'simple_reference', 'simple_speaker_listener', 'fullobs_collect_treasure')


def _scenario_class(name):
    class Scenario(object):
        scenario_name = name

        def make_world(self, num_agents=None):
            if num_agents is not None and name != 'simple_spread':
                raise TypeError('make_world() of %s takes no num_agents' % name)
            return World(name, num_agents)

        # The kernels implement these; the bound methods only identify themselves to MultiAgentEnv.
        def reset_world(self, world):
            raise RuntimeError('reset_world runs on the GPU: call env.reset()')

        def reward(self, agent, world):
            raise RuntimeError('reward runs on the GPU: returned by env.step()')

        def observation(self, agent, world):
            raise RuntimeError('observation runs on the GPU: returned by env.reset()/env.step()')

        def benchmark_data(self, agent, world):
            raise RuntimeError('benchmark_data runs on the GPU: returned in info_n by env.step()')

    if name == 'fullobs_collect_treasure':  # make_env looks for the hook with hasattr (experiments/scenarios.py:174)
        def post_step(self, world):
            raise RuntimeError('post_step runs on the GPU: part of env.step()')
        Scenario.post_step = post_step

    Scenario.__name__ = 'Scenario'
    return Scenario


class _ScenarioModule(object):
    def __init__(self, name):
        self.__name__ = 'multiagent.scenarios.' + name
        self.Scenario = _scenario_class(name)


def load(name):
    if name.endswith('.py'):
        name = name[:-3]
    if name not in SUPPORTED:
        raise ImportError('scenario %r has no sm_100a kernel (supported: %s)' % (name, ', '.join(SUPPORTED)))
    return _ScenarioModule(name)
