"""Loop/object-style float64 restatement of the multi-agent particle environment.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the algorithm the
reference *runs* but does not *contain*: ``experiments/scenarios.py:2-3`` imports
it from the third-party ``multiagent`` package (OpenAI
multi-agent-particle-envs; the MAAC fork adds ``post_step_callback`` /
``discrete_action`` kwargs, used at ``experiments/scenarios.py:179-190``).  No
version is pinned and no copy exists in /root/reference, so this file restates
the published upstream algorithm, keeping upstream's function names
(``World.step``, ``apply_action_force``, ``apply_environment_force``,
``get_collision_force``, ``integrate_state``, ``update_agent_state``,
``MultiAgentEnv.step/reset/_set_action``) so that it can be diffed against
upstream.  **Parity unpinned** by the reference's own tests (it has none).

What IS taken from the reference tree:
  * the three partial-observation functions - experiments/scenarios.py:6-20
    (simple_spread), :23-42 (simple_reference), :45-63 (simple_speaker_listener);
  * the construction flags of ``make_env`` - experiments/scenarios.py:124-192:
    ``world.collaborative = False`` (:171), ``force_discrete_action = True``
    (:191), ``discrete_action=True`` (:182-190), ``make_world(num_agents=n)`` (:170).

EXECUTED-REFERENCE vs RESTATED-UPSTREAM (round 2), by line range of this file:
  * ``SimpleSpread.observation`` / ``SimpleReference.observation`` /
    ``SimpleSpeakerListener.observation`` (the three ``def observation`` bodies) and
    ``make_env`` (the last function): restatements of code that IS in the reference
    tree, now PINNED by executing the reference's own functions - the reference's
    ``experiments/scenarios.py`` is byte-compiled into oracle/_ref by
    ``python -m oracle.build_ref`` and run on tests/_stubs/multiagent;
    tests/test_reference_exec.py::test_reference_make_env_equals_the_oracle_bit_for_bit
    requires bit equality of observations, rewards and benchmark flags with this
    file and with the committed fixtures, for all six scenario / team-size files.
  * everything else (``World.*``, ``MultiAgentEnv.*``, ``reset_world`` / ``reward`` /
    ``benchmark_data`` / ``make_world`` of the three scenarios): restated upstream,
    NOT executable from the reference tree - parity unpinned.  The reference's own
    loops (experiments/run.py, run_BIC.py) and Trainers are executed on top of it
    (tests/test_reference_exec.py, tests/test_gpu_reference_loop.py), which pins
    the CALL contract (argument shapes, in-place one-hot rewrite, reset order,
    return types), not the arithmetic.

Version ambiguities (recorded per SURVEY.md section 8c):
  1. OpenAI vs MAAC fork: identical arithmetic for mass 1, accel None, no walls.
  2. ``make_world(num_agents=n)`` is not in stock simple_spread; we assume
     ``num_landmarks = num_agents`` and everything else stock.
  3. simple_speaker_listener: stock action spaces are Discrete(3)/Discrete(5),
     but main.py:57 builds one actor head of width ``action_space[0].n`` for
     both agents.  Build choice: every agent gets Discrete(5); the speaker's
     message is the first ``dim_c`` entries of its action vector.
"""
import numpy as np

# ----------------------------------------------------------------------------
# gym.spaces stand-ins (gym is not installed; only the attributes the reference
# reads are provided: main.py:51-58 reads .shape[0], .n, .high)
# ----------------------------------------------------------------------------


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete(object):
    """upstream multiagent/multi_discrete.py: array of [min, max] pairs."""

    def __init__(self, array_of_param_array):
        self.low = np.array([x[0] for x in array_of_param_array])
        self.high = np.array([x[1] for x in array_of_param_array])
        self.num_discrete_space = self.low.shape[0]

    def __repr__(self):
        return "MultiDiscrete" + str(self.num_discrete_space)


class Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box" + str(self.shape)


# ----------------------------------------------------------------------------
# multiagent/core.py
# ----------------------------------------------------------------------------


class EntityState(object):
    def __init__(self):
        self.p_pos = None
        self.p_vel = None


class AgentState(EntityState):
    def __init__(self):
        super(AgentState, self).__init__()
        self.c = None


class Action(object):
    def __init__(self):
        self.u = None
        self.c = None


class Entity(object):
    def __init__(self):
        self.name = ''
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.density = 25.0
        self.color = None
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    pass


class Agent(Entity):
    def __init__(self):
        super(Agent, self).__init__()
        self.movable = True
        self.silent = False
        self.blind = False
        self.u_noise = None
        self.c_noise = None
        self.u_range = 1.0
        self.state = AgentState()
        self.action = Action()
        self.action_callback = None


class World(object):
    def __init__(self):
        self.agents = []
        self.landmarks = []
        self.dim_c = 0
        self.dim_p = 2
        self.dim_color = 3
        self.dt = 0.1
        self.damping = 0.25
        self.contact_force = 1e+2
        self.contact_margin = 1e-3

    @property
    def entities(self):
        return self.agents + self.landmarks

    @property
    def policy_agents(self):
        return [a for a in self.agents if a.action_callback is None]

    @property
    def scripted_agents(self):
        return [a for a in self.agents if a.action_callback is not None]

    def step(self):
        for agent in self.scripted_agents:
            agent.action = agent.action_callback(agent, self)
        p_force = [None] * len(self.entities)
        p_force = self.apply_action_force(p_force)
        p_force = self.apply_environment_force(p_force)
        self.integrate_state(p_force)
        for agent in self.agents:
            self.update_agent_state(agent)

    def apply_action_force(self, p_force):
        for i, agent in enumerate(self.agents):
            if agent.movable:
                noise = np.random.randn(*agent.action.u.shape) * agent.u_noise if agent.u_noise else 0.0
                p_force[i] = agent.action.u + noise
        return p_force

    def apply_environment_force(self, p_force):
        ents = self.entities
        for a in range(len(ents)):
            for b in range(a + 1, len(ents)):
                f_a, f_b = self.get_collision_force(ents[a], ents[b])
                if f_a is not None:
                    if p_force[a] is None:
                        p_force[a] = 0.0
                    p_force[a] = f_a + p_force[a]
                if f_b is not None:
                    if p_force[b] is None:
                        p_force[b] = 0.0
                    p_force[b] = f_b + p_force[b]
        return p_force

    def integrate_state(self, p_force):
        for i, entity in enumerate(self.entities):
            if not entity.movable:
                continue
            entity.state.p_vel = entity.state.p_vel * (1 - self.damping)
            if p_force[i] is not None:
                entity.state.p_vel += (p_force[i] / entity.mass) * self.dt
            if entity.max_speed is not None:
                speed = np.sqrt(np.square(entity.state.p_vel[0]) + np.square(entity.state.p_vel[1]))
                if speed > entity.max_speed:
                    entity.state.p_vel = entity.state.p_vel / np.sqrt(
                        np.square(entity.state.p_vel[0]) + np.square(entity.state.p_vel[1])) * entity.max_speed
            entity.state.p_pos += entity.state.p_vel * self.dt

    def update_agent_state(self, agent):
        if agent.silent:
            agent.state.c = np.zeros(self.dim_c)
        else:
            noise = np.random.randn(*agent.action.c.shape) * agent.c_noise if agent.c_noise else 0.0
            agent.state.c = agent.action.c + noise

    def get_collision_force(self, entity_a, entity_b):
        if (not entity_a.collide) or (not entity_b.collide):
            return [None, None]
        if entity_a is entity_b:
            return [None, None]
        delta_pos = entity_a.state.p_pos - entity_b.state.p_pos
        dist = np.sqrt(np.sum(np.square(delta_pos)))
        dist_min = entity_a.size + entity_b.size
        k = self.contact_margin
        penetration = np.logaddexp(0, -(dist - dist_min) / k) * k
        force = self.contact_force * delta_pos / dist * penetration
        force_a = +force if entity_a.movable else None
        force_b = -force if entity_b.movable else None
        return [force_a, force_b]


# ----------------------------------------------------------------------------
# multiagent/scenarios/simple_{spread,reference,speaker_listener}.py
# reset_world draws from the GLOBAL numpy RNG, agents first, then landmarks.
# ----------------------------------------------------------------------------


class BaseScenario(object):
    name = None


def _dist(pa, pb):
    return np.sqrt(np.sum(np.square(pa - pb)))


class SimpleSpread(BaseScenario):
    name = 'simple_spread'

    def make_world(self, num_agents=3):
        world = World()
        world.dim_c = 2
        num_landmarks = num_agents  # ambiguity (2): L = N
        world.collaborative = True
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.15
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.color = np.array([0.35, 0.35, 0.85])
        for landmark in world.landmarks:
            landmark.color = np.array([0.25, 0.25, 0.25])
        for agent in world.agents:
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def is_collision(self, agent1, agent2):
        dist = _dist(agent1.state.p_pos, agent2.state.p_pos)
        dist_min = agent1.size + agent2.size
        return True if dist < dist_min else False

    def benchmark_data(self, agent, world):
        rew = 0
        collisions = 0
        occupied_landmarks = 0
        min_dists = 0
        for l in world.landmarks:
            dists = [_dist(a.state.p_pos, l.state.p_pos) for a in world.agents]
            min_dists += min(dists)
            rew -= min(dists)
            if min(dists) < 0.1:
                occupied_landmarks += 1
        if agent.collide:
            for a in world.agents:
                if self.is_collision(a, agent):
                    rew -= 1
                    collisions += 1
        return (rew, collisions, min_dists, occupied_landmarks)

    def reward(self, agent, world):
        rew = 0
        for l in world.landmarks:
            dists = [_dist(a.state.p_pos, l.state.p_pos) for a in world.agents]
            rew -= min(dists)
        if agent.collide:
            for a in world.agents:  # includes a is agent: constant -1
                if self.is_collision(a, agent):
                    rew -= 1
        return rew

    def observation(self, agent, world):
        """experiments/scenarios.py:6-20 (local_obs_simple_spread)."""
        entity_pos = []
        for entity in world.landmarks:
            entity_pos.append(entity.state.p_pos - agent.state.p_pos)
        return np.concatenate([agent.state.p_vel] + [agent.state.p_pos] + entity_pos)


class SimpleReference(BaseScenario):
    name = 'simple_reference'

    def make_world(self):
        world = World()
        world.dim_c = 10
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
        world.landmarks = [Landmark() for _ in range(3)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.goal_a = None
            agent.goal_b = None
        world.agents[0].goal_a = world.agents[1]
        world.agents[0].goal_b = np.random.choice(world.landmarks)
        world.agents[1].goal_a = world.agents[0]
        world.agents[1].goal_b = np.random.choice(world.landmarks)
        for agent in world.agents:
            agent.color = np.array([0.25, 0.25, 0.25])
        world.landmarks[0].color = np.array([0.75, 0.25, 0.25])
        world.landmarks[1].color = np.array([0.25, 0.75, 0.25])
        world.landmarks[2].color = np.array([0.25, 0.25, 0.75])
        world.agents[0].goal_a.color = world.agents[0].goal_b.color
        world.agents[1].goal_a.color = world.agents[1].goal_b.color
        for agent in world.agents:
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def reward(self, agent, world):
        if agent.goal_a is None or agent.goal_b is None:
            return 0.0
        dist2 = np.sum(np.square(agent.goal_a.state.p_pos - agent.goal_b.state.p_pos))
        return -dist2

    def benchmark_data(self, agent, world):
        return self.reward(agent, world)

    def observation(self, agent, world):
        """experiments/scenarios.py:23-42 (local_obs_simple_reference)."""
        goal_color = [np.zeros(world.dim_color), np.zeros(world.dim_color)]
        if agent.goal_b is not None:
            goal_color[1] = agent.goal_b.color
        entity_pos = []
        for entity in world.landmarks:
            entity_pos.append(entity.state.p_pos - agent.state.p_pos)
        comm = []
        for other in world.agents:
            if other is agent:
                continue
            comm.append(other.state.c)
        return np.concatenate([agent.state.p_vel] + entity_pos + [goal_color[1]] + comm)


class SimpleSpeakerListener(BaseScenario):
    name = 'simple_speaker_listener'

    def make_world(self):
        world = World()
        world.dim_c = 3
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.size = 0.075
        world.agents[0].movable = False  # speaker
        world.agents[1].silent = True  # listener
        world.landmarks = [Landmark() for _ in range(3)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.04
        self.reset_world(world)
        return world

    def reset_world(self, world):
        for agent in world.agents:
            agent.goal_a = None
            agent.goal_b = None
        world.agents[0].goal_a = world.agents[1]
        world.agents[0].goal_b = np.random.choice(world.landmarks)
        for agent in world.agents:
            agent.color = np.array([0.25, 0.25, 0.25])
        world.landmarks[0].color = np.array([0.65, 0.15, 0.15])
        world.landmarks[1].color = np.array([0.15, 0.65, 0.15])
        world.landmarks[2].color = np.array([0.15, 0.15, 0.65])
        world.agents[0].goal_a.color = world.agents[0].goal_b.color + np.array([0.45, 0.45, 0.45])
        for agent in world.agents:
            agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
        for landmark in world.landmarks:
            landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            landmark.state.p_vel = np.zeros(world.dim_p)

    def reward(self, agent, world):
        a = world.agents[0]
        dist2 = np.sum(np.square(a.goal_a.state.p_pos - a.goal_b.state.p_pos))
        return -dist2

    def benchmark_data(self, agent, world):
        return self.reward(agent, world)

    def observation(self, agent, world):
        """experiments/scenarios.py:45-63 (local_obs_simple_speaker_listener):
        the comm list is built and then dropped, both agents get 11 floats."""
        goal_color = np.zeros(world.dim_color)
        if agent.goal_b is not None:
            goal_color = agent.goal_b.color
        entity_pos = []
        for entity in world.landmarks:
            entity_pos.append(entity.state.p_pos - agent.state.p_pos)
        return np.concatenate([agent.state.p_vel] + entity_pos + [goal_color])


SCENARIOS = {
    'simple_spread': SimpleSpread,
    'simple_reference': SimpleReference,
    'simple_speaker_listener': SimpleSpeakerListener,
}


# ----------------------------------------------------------------------------
# multiagent/environment.py
# ----------------------------------------------------------------------------


class MultiAgentEnv(object):
    def __init__(self, world, reset_callback=None, reward_callback=None,
                 observation_callback=None, info_callback=None,
                 done_callback=None, post_step_callback=None,
                 shared_viewer=True, discrete_action=True, uniform_action_width=None):
        self.world = world
        self.agents = self.world.policy_agents
        self.n = len(world.agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.post_step_callback = post_step_callback
        self.discrete_action_space = discrete_action
        self.discrete_action_input = False
        self.force_discrete_action = world.discrete_action if hasattr(world, 'discrete_action') else False
        self.shared_reward = world.collaborative if hasattr(world, 'collaborative') else False
        self.time = 0
        # ambiguity (3): when set, every agent's action vector has this many
        # leading entries reserved for its (single) physical/comm head.
        self.uniform_action_width = uniform_action_width

        self.action_space = []
        self.observation_space = []
        for agent in self.agents:
            total_action_space = []
            u_action_space = Discrete(world.dim_p * 2 + 1)
            if agent.movable:
                total_action_space.append(u_action_space)
            c_action_space = Discrete(world.dim_c)
            if not agent.silent:
                total_action_space.append(c_action_space)
            if len(total_action_space) > 1:
                act_space = MultiDiscrete([[0, s.n - 1] for s in total_action_space])
                self.action_space.append(act_space)
            elif uniform_action_width is not None:
                self.action_space.append(Discrete(uniform_action_width))
            else:
                self.action_space.append(total_action_space[0])
            obs_dim = len(observation_callback(agent, self.world))
            self.observation_space.append(Box(low=-np.inf, high=+np.inf, shape=(obs_dim,), dtype=np.float32))
            agent.action.c = np.zeros(self.world.dim_c)

    def seed(self, seed=None):
        np.random.seed(1 if seed is None else seed)

    def step(self, action_n):
        obs_n, reward_n, done_n = [], [], []
        info_n = {'n': []}
        self.agents = self.world.policy_agents
        for i, agent in enumerate(self.agents):
            self._set_action(action_n[i], agent, self.action_space[i])
        self.world.step()
        for agent in self.agents:
            obs_n.append(self._get_obs(agent))
            reward_n.append(self._get_reward(agent))
            done_n.append(self._get_done(agent))
            info_n['n'].append(self._get_info(agent))
        reward = np.sum(reward_n)
        if self.shared_reward:
            reward_n = [reward] * self.n
        # MAAC fork: the scenario hook runs after the step's observations and rewards were taken (None for the
        # three stock scenarios; oracle/maac_ref.py, ambiguity 1, for fullobs_collect_treasure)
        if self.post_step_callback is not None:
            self.post_step_callback(self.world)
        return obs_n, reward_n, done_n, info_n

    def reset(self):
        self.reset_callback(self.world)
        obs_n = []
        self.agents = self.world.policy_agents
        for agent in self.agents:
            obs_n.append(self._get_obs(agent))
        return obs_n

    def _get_info(self, agent):
        if self.info_callback is None:
            return {}
        return self.info_callback(agent, self.world)

    def _get_obs(self, agent):
        if self.observation_callback is None:
            return np.zeros(0)
        return self.observation_callback(agent, self.world)

    def _get_done(self, agent):
        if self.done_callback is None:
            return False
        return self.done_callback(agent, self.world)

    def _get_reward(self, agent):
        if self.reward_callback is None:
            return 0.0
        return self.reward_callback(agent, self.world)

    def _set_action(self, action, agent, action_space, time=None):
        agent.action.u = np.zeros(self.world.dim_p)
        agent.action.c = np.zeros(self.world.dim_c)
        if isinstance(action_space, MultiDiscrete):
            act = []
            size = action_space.high - action_space.low + 1
            index = 0
            for s in size:
                act.append(action[index:(index + s)])
                index += s
            action = act
        else:
            action = [action]

        if agent.movable:
            if self.discrete_action_input:
                agent.action.u = np.zeros(self.world.dim_p)
                if action[0] == 1: agent.action.u[0] = -1.0
                if action[0] == 2: agent.action.u[0] = +1.0
                if action[0] == 3: agent.action.u[1] = -1.0
                if action[0] == 4: agent.action.u[1] = +1.0
            else:
                if self.force_discrete_action:
                    d = np.argmax(action[0])
                    action[0][:] = 0.0
                    action[0][d] = 1.0
                if self.discrete_action_space:
                    agent.action.u[0] += action[0][1] - action[0][2]
                    agent.action.u[1] += action[0][3] - action[0][4]
                else:
                    agent.action.u = action[0]
            sensitivity = 5.0
            if agent.accel is not None:
                sensitivity = agent.accel
            agent.action.u *= sensitivity
            action = action[1:]
        if not agent.silent:
            if self.discrete_action_input:
                agent.action.c = np.zeros(self.world.dim_c)
                agent.action.c[action[0]] = 1.0
            else:
                # ambiguity (3): a uniform-width head is cut to dim_c
                agent.action.c = np.asarray(action[0], dtype=np.float64)[:self.world.dim_c]
            action = action[1:]
        assert len(action) == 0


def make_env(scenario_name, n=None, local_observation=True, benchmark=False, discrete_action=True):
    """Restates experiments/scenarios.py:124-192 on top of the classes above.

    ``local_observation`` is accepted for signature parity; the scenario
    classes here always use the reference's partial observations (that is the
    only branch main.py:39 and main_scalability_1.py:36 take).
    """
    if scenario_name == 'fullobs_collect_treasure':
        from . import maac_ref  # noqa: F401  (registers the MAAC-fork scenario)
    if scenario_name not in SCENARIOS:
        raise ValueError('unsupported scenario: %r' % (scenario_name,))
    scenario = SCENARIOS[scenario_name]()
    if n is None:
        world = scenario.make_world()
    else:
        world = scenario.make_world(num_agents=n)
    world.collaborative = False  # scenarios.py:171
    uniform = 5 if scenario_name == 'simple_speaker_listener' else None
    env = MultiAgentEnv(world, reset_callback=scenario.reset_world,
                        reward_callback=scenario.reward,
                        observation_callback=scenario.observation,
                        post_step_callback=getattr(scenario, 'post_step', None),  # scenarios.py:174-177
                        info_callback=scenario.benchmark_data if benchmark else None,
                        discrete_action=discrete_action,
                        uniform_action_width=uniform)
    env.force_discrete_action = True  # scenarios.py:191
    env.scenario = scenario
    return env


# ----------------------------------------------------------------------------
# state injection helpers used by the parity tests
# ----------------------------------------------------------------------------


def set_state(env, agent_pos, agent_vel, landmark_pos, goals=None):
    """agent_pos/vel [N,2], landmark_pos [L,2], goals: landmark index per agent (-1 = None)."""
    w = env.world
    for i, a in enumerate(w.agents):
        a.state.p_pos = np.array(agent_pos[i], dtype=np.float64)
        a.state.p_vel = np.array(agent_vel[i], dtype=np.float64)
        a.state.c = np.zeros(w.dim_c)
    for i, l in enumerate(w.landmarks):
        l.state.p_pos = np.array(landmark_pos[i], dtype=np.float64)
        l.state.p_vel = np.zeros(w.dim_p)
    if goals is not None:
        for i, a in enumerate(w.agents):
            g = int(goals[i])
            if hasattr(a, 'goal_b') or g >= 0:
                a.goal_b = w.landmarks[g] if g >= 0 else None


def get_obs(env):
    return [env._get_obs(a) for a in env.agents]
