"""Float64 loop restatement of the MAAC fork's engine additions and of ``fullobs_collect_treasure``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md section 8f-3: the reference names the scenario at
``main.py:24-25``, patches its observation at ``experiments/scenarios.py:162-163`` with
``local_obs_collect_treasure`` (``experiments/scenarios.py:95-121``, IN the reference tree) and passes the
scenario's ``post_step`` hook to ``MultiAgentEnv`` (``experiments/scenarios.py:174-190``).  Everything else the
scenario needs lives in the MAAC fork of the particle-env package (shariqiqbal2810/multiagent-particle-envs:
``multiagent/core.py`` with ``cache_dists`` / mass ratios / per-agent ``accel``, and
``multiagent/scenarios/fullobs_collect_treasure.py``), which is neither vendored in /root/reference nor
installable here, and no version is pinned anywhere in the tree.

**Parity unpinned** for the fork-side arithmetic: this file restates the published fork from the builder's
knowledge of it, keeping the fork's function names (``World.calculate_distances``, ``apply_action_force``,
``get_entity_collision_force``, ``Scenario.post_step`` / ``collector_reward`` / ``deposit_reward`` /
``global_reward`` / ``is_collision``) so that it can be diffed against the fork later.

What IS pinned by executing reference code: ``CollectTreasure.observation`` below restates
``local_obs_collect_treasure``; tests/test_reference_exec.py runs the reference's own compiled function
(through its ``make_env``) on this world and requires bit equality.

Choices where the fork's text could not be re-checked (each one is a single line below, marked AMBIGUITY):
  1. ``MultiAgentEnv.step`` calls ``post_step_callback`` AFTER the observations and rewards of the step have been
     taken.  (The collecting reward tests ``is_collision(collector, treasure)`` for collectors that hold nothing;
     if the hook ran first the collector would already hold the treasure and the treasure would already sit at
     (-999, -999), so that reward could never fire.)
  2. ``apply_action_force`` multiplies the action by ``mass * accel`` although ``_set_action`` has already scaled it
     by ``accel`` (the fork's double application): acceleration per unit action = accel^2 = 2.25 for every agent.
  3. ``deposit_reward`` without a matching holder: mean offset of ALL seven other agents (``n_visible = 7``).
  4. a treasure that was picked up respawns in the NEXT ``post_step`` (``respawn_prob = 1.0``; the draw order is
     probability, position, type).
"""
import numpy as np

from . import mpe_ref
from .mpe_ref import Agent, BaseScenario, Landmark, World


class ForkWorld(World):
    """multiagent/core.py of the MAAC fork: the stock World plus cached distances, mass-aware contact forces and
    ``mass * accel`` action forces.  With ``cache_dists=False``, unit masses and ``accel=None`` it is the stock
    World (ambiguity (1) of oracle/mpe_ref.py)."""

    def __init__(self):
        super(ForkWorld, self).__init__()
        self.walls = []
        self.cache_dists = False
        self.cached_dist_vect = None
        self.cached_dist_mag = None
        self.min_dists = None

    def calculate_distances(self):
        ents = self.entities
        if self.cached_dist_vect is None:
            self.cached_dist_vect = np.zeros((len(ents), len(ents), self.dim_p))
            self.min_dists = np.zeros((len(ents), len(ents)))
            for ia, entity_a in enumerate(ents):
                for ib in range(ia + 1, len(ents)):
                    min_dist = entity_a.size + ents[ib].size
                    self.min_dists[ia, ib] = min_dist
                    self.min_dists[ib, ia] = min_dist
        for ia, entity_a in enumerate(ents):
            for ib in range(ia + 1, len(ents)):
                delta_pos = entity_a.state.p_pos - ents[ib].state.p_pos
                self.cached_dist_vect[ia, ib, :] = delta_pos
                self.cached_dist_vect[ib, ia, :] = -delta_pos
        self.cached_dist_mag = np.linalg.norm(self.cached_dist_vect, axis=2)
        self.cached_collisions = (self.cached_dist_mag <= self.min_dists)

    def step(self):
        super(ForkWorld, self).step()
        if self.cache_dists:
            self.calculate_distances()

    def apply_action_force(self, p_force):
        for i, agent in enumerate(self.agents):
            if agent.movable:
                noise = np.random.randn(*agent.action.u.shape) * agent.u_noise if agent.u_noise else 0.0
                # AMBIGUITY 2: force = mass * accel * action (the action already carries one factor accel)
                p_force[i] = (agent.mass * agent.accel if agent.accel is not None else agent.mass) * agent.action.u + noise
        return p_force

    def apply_environment_force(self, p_force):
        ents = self.entities
        for a in range(len(ents)):
            for b in range(a + 1, len(ents)):
                f_a, f_b = self.get_entity_collision_force(a, b)
                if f_a is not None:
                    if p_force[a] is None:
                        p_force[a] = 0.0
                    p_force[a] = f_a + p_force[a]
                if f_b is not None:
                    if p_force[b] is None:
                        p_force[b] = 0.0
                    p_force[b] = f_b + p_force[b]
        return p_force

    def get_entity_collision_force(self, ia, ib):
        entity_a, entity_b = self.entities[ia], self.entities[ib]
        if (not entity_a.collide) or (not entity_b.collide):
            return [None, None]
        if (not entity_a.movable) and (not entity_b.movable):
            return [None, None]
        if entity_a is entity_b:
            return [None, None]
        if self.cache_dists:
            delta_pos = self.cached_dist_vect[ia, ib]
            dist = self.cached_dist_mag[ia, ib]
            dist_min = self.min_dists[ia, ib]
        else:
            delta_pos = entity_a.state.p_pos - entity_b.state.p_pos
            dist = np.sqrt(np.sum(np.square(delta_pos)))
            dist_min = entity_a.size + entity_b.size
        k = self.contact_margin
        penetration = np.logaddexp(0, -(dist - dist_min) / k) * k
        force = self.contact_force * delta_pos / dist * penetration
        if entity_a.movable and entity_b.movable:
            force_ratio = entity_b.mass / entity_a.mass  # consider mass in collisions
            force_a = force_ratio * force
            force_b = -(1 / force_ratio) * force
        else:
            force_a = +force if entity_a.movable else None
            force_b = -force if entity_b.movable else None
        return [force_a, force_b]


class NumpyDraws(object):
    """the scenario's random draws from the GLOBAL numpy generator, in the fork's call order"""

    def position(self, world, entity_index, bound):
        return np.random.uniform(low=-bound, high=bound, size=world.dim_p)

    def treasure_type(self, world, treasure_index):
        return np.random.choice(world.treasure_types)

    def respawn(self, world, treasure_index, prob):
        return np.random.uniform() <= prob


class PhiloxDraws(object):
    """the kernels' counter-based streams (oracle/philox.py) behind the same three calls, for one env: the test
    sets ``episode`` / ``tstep`` to the env's counters before it steps the oracle"""

    def __init__(self, seed, gid):
        from . import philox
        self.philox, self.seed, self.gid = philox, seed, gid
        self.episode, self.tstep = 0, 0
        self._resp = {}

    def _reset(self):
        a, t, ty = self.philox.treasure_reset(self.seed, np.array([self.gid]), self.episode)
        return a[0], t[0], ty[0]

    def _respawn(self, l):
        p, ty = self.philox.treasure_respawn(self.seed, np.array([self.gid]), self.episode, self.tstep, l)
        return p[0], int(ty[0])

    def position(self, world, entity_index, bound):
        if bound == 1.0:
            return self._reset()[0][entity_index].copy()
        l = entity_index - len(world.agents)
        if self._resp.get(l):  # inside post_step: the respawn stream
            self._resp[l] = False
            return self._respawn(l)[0].copy()
        return self._reset()[1][l].copy()

    def treasure_type(self, world, treasure_index):
        if self._resp.get(treasure_index) is False:
            del self._resp[treasure_index]
            return self._respawn(treasure_index)[1]
        return int(self._reset()[2][treasure_index])

    def respawn(self, world, treasure_index, prob):
        self._resp[treasure_index] = True
        return True


class CollectTreasure(BaseScenario):
    """multiagent/scenarios/fullobs_collect_treasure.py: 6 collectors + 2 deposits (all agents), 6 treasures."""
    name = 'fullobs_collect_treasure'
    NUM_AGENTS, NUM_COLLECTORS = 8, 6

    def __init__(self):
        self.draws = NumpyDraws()  # tests substitute the kernels' Philox streams (tests/_treasure.py)

    def make_world(self):
        world = ForkWorld()
        world.cache_dists = True
        world.dim_c = 2
        num_agents, num_collectors = self.NUM_AGENTS, self.NUM_COLLECTORS
        num_deposits = num_agents - num_collectors
        world.treasure_types = list(range(num_deposits))
        num_treasures = num_collectors
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.i = i
            agent.name = 'agent %d' % i
            agent.collector = True if i < num_collectors else False
            if not agent.collector:
                agent.d_i = i - num_collectors
            agent.collide = True
            agent.silent = True
            agent.ghost = True
            agent.holding = None
            agent.size = 0.05 if agent.collector else 0.075
            agent.accel = 1.5
            agent.initial_mass = 1.0 if agent.collector else 2.25
            agent.max_speed = 1.0
        world.landmarks = [Landmark() for _ in range(num_treasures)]
        for i, landmark in enumerate(world.landmarks):
            landmark.i = i + num_agents
            landmark.name = 'treasure %d' % i
            landmark.respawn_prob = 1.0
            landmark.type = self.draws.treasure_type(world, i)
            landmark.alive = True
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.025
            landmark.boundary = False
        world.walls = []
        self.reset_world(world)
        self.reset_cached_rewards()
        return world

    def collectors(self, world):
        return [a for a in world.agents if a.collector]

    def deposits(self, world):
        return [a for a in world.agents if not a.collector]

    def treasures(self, world):
        return world.landmarks

    def reset_cached_rewards(self):
        self.global_collecting_reward = None
        self.global_holding_reward = None
        self.global_deposit_reward = None

    def post_step(self, world):
        self.reset_cached_rewards()
        for li, l in enumerate(self.treasures(world)):
            if l.alive:
                for a in self.collectors(world):
                    if a.holding is None and self.is_collision(l, a, world):
                        l.alive = False
                        a.holding = l.type
                        l.state.p_pos = np.array([-999., -999.])
                        break
            else:
                if self.draws.respawn(world, li, l.respawn_prob):  # AMBIGUITY 4
                    bound = 0.95
                    l.state.p_pos = self.draws.position(world, l.i, bound)
                    l.type = self.draws.treasure_type(world, li)
                    l.alive = True
        for a in self.collectors(world):
            if a.holding is not None:
                for d in self.deposits(world):
                    if d.d_i == a.holding and self.is_collision(a, d, world):
                        a.holding = None

    def reset_world(self, world):
        for i, agent in enumerate(world.agents):
            agent.state.p_pos = self.draws.position(world, i, 1.0)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
            agent.holding = None
        for i, landmark in enumerate(world.landmarks):
            bound = 0.95
            landmark.type = self.draws.treasure_type(world, i)
            landmark.state.p_pos = self.draws.position(world, landmark.i, bound)
            landmark.state.p_vel = np.zeros(world.dim_p)
            landmark.alive = True
        world.calculate_distances()

    def benchmark_data(self, agent, world):
        if agent.collector:
            if agent.holding is not None:
                for d in self.deposits(world):
                    if d.d_i == agent.holding and self.is_collision(d, agent, world):
                        return 1
            else:
                for t in self.treasures(world):
                    if self.is_collision(t, agent, world):
                        return 1
        return 0

    def is_collision(self, agent1, agent2, world):
        dist = world.cached_dist_mag[agent1.i, agent2.i]
        dist_min = agent1.size + agent2.size
        return True if dist < dist_min else False

    def reward(self, agent, world):
        return self.collector_reward(agent, world) if agent.collector else self.deposit_reward(agent, world)

    def deposit_reward(self, agent, world):
        rew = 0
        # shaped: distance to the closest collector that holds this deposit's type, else mean offset of the others
        dists_to_holding = [world.cached_dist_mag[agent.i, a.i] for a in self.collectors(world)
                            if a.holding == agent.d_i]
        if len(dists_to_holding) > 0:
            rew -= 0.1 * min(dists_to_holding)
        else:
            n_visible = 7
            other_agent_inds = [a.i for a in world.agents if a is not agent]  # AMBIGUITY 3
            closest_agents = sorted(zip(world.cached_dist_mag[other_agent_inds, agent.i], other_agent_inds))[:n_visible]
            closest_inds = list(i for _, i in closest_agents)
            closest_avg_dist_vect = world.cached_dist_vect[closest_inds, agent.i].mean(axis=0)
            rew -= 0.1 * np.linalg.norm(closest_avg_dist_vect)
        rew += self.global_reward(world)
        return rew

    def collector_reward(self, agent, world):
        rew = 0
        # penalize collisions between collectors
        rew -= 5 * sum(self.is_collision(agent, a, world) for a in self.collectors(world) if a is not agent)
        if agent.holding is None:
            rew -= 0.1 * min(world.cached_dist_mag[t.i, agent.i] for t in self.treasures(world))
        else:
            rew -= 0.1 * min(world.cached_dist_mag[d.i, agent.i] for d in self.deposits(world)
                             if d.d_i == agent.holding)
        rew += self.global_reward(world)
        return rew

    def global_reward(self, world):
        if self.global_deposit_reward is None:
            self.calc_global_deposit_reward(world)
        if self.global_collecting_reward is None:
            self.calc_global_collecting_reward(world)
        return self.global_deposit_reward + self.global_collecting_reward

    def calc_global_collecting_reward(self, world):
        rew = 0
        for t in self.treasures(world):
            rew += 5 * sum(self.is_collision(a, t, world) for a in self.collectors(world) if a.holding is None)
        self.global_collecting_reward = rew

    def calc_global_deposit_reward(self, world):
        rew = 0
        for d in self.deposits(world):
            rew += 5 * sum(self.is_collision(d, a, world) for a in self.collectors(world) if a.holding == d.d_i)
        self.global_deposit_reward = rew

    def get_agent_encoding(self, agent, world):
        encoding = []
        n_treasure_types = len(world.treasure_types)
        if agent.collector:
            encoding.append(np.zeros(n_treasure_types))
            encoding.append((np.arange(n_treasure_types) == agent.holding))
        else:
            encoding.append((np.arange(n_treasure_types) == agent.d_i))
            encoding.append(np.zeros(n_treasure_types))
        return np.concatenate(encoding)

    def stock_observation(self, agent, world, n_visible=7):
        """the fork's own (full) observation: 7 nearest agents with velocity + encoding, 7 nearest treasures"""
        other_agents = [a.i for a in world.agents if a is not agent]
        closest_agents = sorted(zip(world.cached_dist_mag[other_agents, agent.i], other_agents))[:n_visible]
        treasures = [t.i for t in self.treasures(world)]
        closest_treasures = sorted(zip(world.cached_dist_mag[treasures, agent.i], treasures))[:n_visible]
        n_treasure_types = len(world.treasure_types)
        obs = [agent.state.p_pos, agent.state.p_vel]
        if agent.collector:
            obs.append((np.arange(n_treasure_types) == agent.holding))
        for _, i in closest_agents:
            a = world.entities[i]
            obs.append(world.cached_dist_vect[i, agent.i])
            obs.append(a.state.p_vel)
            obs.append(self.get_agent_encoding(a, world))
        for _, i in closest_treasures:
            t = world.entities[i]
            obs.append(world.cached_dist_vect[i, agent.i])
            obs.append((np.arange(n_treasure_types) == t.type))
        return np.concatenate(obs)

    def observation(self, agent, world):
        """experiments/scenarios.py:95-121 (local_obs_collect_treasure): own position, velocity, holding one-hot
        (for every agent), then the six treasures nearest first (ties by index) as (offset, type one-hot) - no other
        agent is visible (``n_visible = 0``).  30 floats."""
        treasures = [t.i for t in self.treasures(world)]
        closest_treasures = sorted(zip(world.cached_dist_mag[treasures, agent.i], treasures))[:7]
        n_treasure_types = len(world.treasure_types)
        obs = [agent.state.p_pos, agent.state.p_vel]
        obs.append((np.arange(n_treasure_types) == agent.holding))
        for _, i in closest_treasures:
            t = world.entities[i]
            obs.append(world.cached_dist_vect[i, agent.i])
            obs.append((np.arange(n_treasure_types) == t.type))
        return np.concatenate(obs)


mpe_ref.SCENARIOS['fullobs_collect_treasure'] = CollectTreasure


# ----------------------------------------------------------------------------
# state injection / readback used by the parity tests
# ----------------------------------------------------------------------------

def pack_flags(types, alive, holding):
    """the kernels' per-env state word: bit l = type of treasure l, bit 6 + l = alive, bits 12 + 2 i .. = holding of
    collector i (0 = nothing, 1 + type otherwise)"""
    w = 0
    for l in range(6):
        w |= (int(types[l]) & 1) << l
        w |= (1 if alive[l] else 0) << (6 + l)
    for i in range(6):
        h = holding[i]
        w |= (0 if h is None or h < 0 else 1 + int(h)) << (12 + 2 * i)
    return w


def unpack_flags(w):
    w = int(w)
    types = [(w >> l) & 1 for l in range(6)]
    alive = [bool((w >> (6 + l)) & 1) for l in range(6)]
    holding = [((w >> (12 + 2 * i)) & 3) - 1 for i in range(6)]  # -1 = nothing
    return types, alive, holding


def set_state(env, agent_pos, agent_vel, treasure_pos, flags):
    w = env.world
    types, alive, holding = unpack_flags(flags)
    for i, a in enumerate(w.agents):
        a.state.p_pos = np.array(agent_pos[i], dtype=np.float64)
        a.state.p_vel = np.array(agent_vel[i], dtype=np.float64)
        a.state.c = np.zeros(w.dim_c)
        if a.collector:
            a.holding = None if holding[i] < 0 else holding[i]
    for l, t in enumerate(w.landmarks):
        t.state.p_pos = np.array(treasure_pos[l], dtype=np.float64)
        t.type = types[l]
        t.alive = alive[l]
    w.calculate_distances()
    env.scenario.reset_cached_rewards()


def get_flags(env):
    w = env.world
    return pack_flags([t.type for t in w.landmarks], [t.alive for t in w.landmarks],
                      [a.holding for a in w.agents[:6]])
