"""upstream multiagent/scenarios/simple_speaker_listener.py with its STOCK observation (the speaker sees only the
goal colour, the listener its velocity, the landmarks and the message)."""
import numpy as np

from oracle import mpe_ref


class Scenario(mpe_ref.SimpleSpeakerListener):
    def observation(self, agent, world):
        goal_color = np.zeros(world.dim_color)
        if agent.goal_b is not None:
            goal_color = agent.goal_b.color
        entity_pos = [entity.state.p_pos - agent.state.p_pos for entity in world.landmarks]
        comm = [other.state.c for other in world.agents if other is not agent and other.state.c is not None]
        if not agent.movable:
            return np.concatenate([goal_color])
        if agent.silent:
            return np.concatenate([agent.state.p_vel] + entity_pos + comm)
