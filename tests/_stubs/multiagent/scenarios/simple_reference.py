"""upstream multiagent/scenarios/simple_reference.py with its STOCK observation."""
import numpy as np

from oracle import mpe_ref


class Scenario(mpe_ref.SimpleReference):
    def observation(self, agent, world):
        goal_color = [np.zeros(world.dim_color), np.zeros(world.dim_color)]
        if agent.goal_b is not None:
            goal_color[1] = agent.goal_b.color
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        comm = [other.state.c for other in world.agents if other is not agent]
        return np.concatenate([agent.state.p_vel] + entity_pos + [goal_color[1]] + comm)
