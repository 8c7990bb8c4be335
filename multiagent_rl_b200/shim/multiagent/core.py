"""``multiagent.core``: the world descriptor ``Scenario.make_world`` returns.  It carries no state - entity state
lives in HBM - only what experiments/scenarios.py:168-171 and env construction read or write."""


class World(object):
    def __init__(self, scenario_name, num_agents=None):
        self.scenario_name = scenario_name
        self.num_agents = num_agents
        self.collaborative = True   # stock scenarios set True; experiments/scenarios.py:171 overwrites with False
        self.dim_p = 2
        self.dim_color = 3
        self.dim_c = {'simple_spread': 2, 'simple_reference': 10, 'simple_speaker_listener': 3,
                      'fullobs_collect_treasure': 2}[scenario_name]
