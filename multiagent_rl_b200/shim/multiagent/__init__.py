"""Drop-in ``multiagent`` package backed by libmpe_b200.so (see multiagent_rl_b200/shim/__init__.py)."""
