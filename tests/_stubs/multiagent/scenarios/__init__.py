"""``multiagent.scenarios`` - upstream's ``load(name)`` returns the scenario script as a module."""
import importlib


def load(name):
    if name.endswith('.py'):
        name = name[:-3]
    return importlib.import_module(__name__ + '.' + name)
