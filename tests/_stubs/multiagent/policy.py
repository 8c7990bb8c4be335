"""``multiagent.policy`` - upstream's Policy base class (test_env/custom_policy.py:5 subclasses it)."""


class Policy(object):
    def __init__(self):
        pass

    def action(self, obs):
        raise NotImplementedError()
