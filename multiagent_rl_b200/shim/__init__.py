"""``multiagent`` shim: makes the reference's UNCHANGED ``experiments/scenarios.py`` build the CUDA environment.

The reference constructs its env through the third-party ``multiagent`` package
(experiments/scenarios.py:2-3,150,168-190).  Putting this directory first on ``sys.path``

    import multiagent_rl_b200.shim as shim; shim.install()      # or PYTHONPATH=<repo>/multiagent_rl_b200/shim
    from experiments.scenarios import make_env                  # the reference's own file, untouched

resolves ``multiagent.scenarios.load(...)`` / ``multiagent.environment.MultiAgentEnv(...)`` to thin descriptors
whose only job is to hand the scenario name and team size to ``BatchedMultiAgentEnv``; all physics, observations
and rewards run in libmpe_b200.so.  The callbacks the reference passes are not called (they would be Python on
the hot path); they are inspected so that an unsupported combination fails loudly instead of silently computing
something else.
"""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    """Put the shim's ``multiagent`` package first on sys.path (idempotent)."""
    if SHIM_DIR in sys.path:
        sys.path.remove(SHIM_DIR)
    sys.path.insert(0, SHIM_DIR)
    return SHIM_DIR
