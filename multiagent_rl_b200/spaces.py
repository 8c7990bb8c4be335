"""Stand-ins for the gym spaces the reference reads off the env (gym is not a dependency):
main.py:51-58 uses ``observation_space[0].shape[0]``, ``action_space[0].n`` and
``action_space[0].high`` (+1, ``.tolist()``)."""
import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return 'Discrete(%d)' % self.n


class MultiDiscrete(object):
    """upstream multiagent/multi_discrete.py: one [min, max] pair per sub-space."""

    def __init__(self, array_of_param_array):
        self.low = np.array([x[0] for x in array_of_param_array])
        self.high = np.array([x[1] for x in array_of_param_array])
        self.num_discrete_space = self.low.shape[0]
        self.shape = (self.num_discrete_space,)

    def sample(self):
        r = np.random.rand(self.num_discrete_space)
        return [int(x) for x in np.floor(np.multiply((self.high - self.low + 1.), r) + self.low)]

    def __repr__(self):
        return 'MultiDiscrete' + str(self.num_discrete_space)


class Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return 'Box' + str(self.shape)
