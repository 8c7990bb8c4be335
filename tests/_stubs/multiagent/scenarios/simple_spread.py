"""upstream multiagent/scenarios/simple_spread.py with its STOCK observation (everything visible)."""
import numpy as np

from oracle import mpe_ref


class Scenario(mpe_ref.SimpleSpread):
    def observation(self, agent, world):
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        comm, other_pos = [], []
        for other in world.agents:
            if other is agent:
                continue
            comm.append(other.state.c)
            other_pos.append(other.state.p_pos - agent.state.p_pos)
        return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos + other_pos + comm)
