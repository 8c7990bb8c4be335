"""MAAC fork multiagent/scenarios/fullobs_collect_treasure.py with its STOCK observation (7 nearest agents with
velocity and encoding, 7 nearest treasures; collectors only carry the holding one-hot), so that the 30-wide partial
observation the tests see can only come from the reference's patch (experiments/scenarios.py:162-163)."""
from oracle import maac_ref


class Scenario(maac_ref.CollectTreasure):
    def observation(self, agent, world):
        return self.stock_observation(agent, world)
