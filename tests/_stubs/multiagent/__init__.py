"""Stand-in for the third-party ``multiagent`` package (OpenAI multi-agent-particle-envs lineage) that the
reference imports at experiments/scenarios.py:2-3 and that is neither vendored nor installable here.

TEST INFRASTRUCTURE.  It exists so that the reference's OWN ``experiments/scenarios.py`` (``make_env`` and the
``local_obs_*`` functions, compiled into oracle/_ref by ``python -m oracle.build_ref``) can be imported and
EXECUTED: engine classes come from the oracle's restatement (oracle/mpe_ref.py: ``core`` / ``environment``);
the ``scenarios`` sub-package holds Scenario classes with upstream's STOCK full observations, so the partial
observations the tests see can only come from the reference's monkey-patch (scenarios.py:151-164).

What a test through this package pins, and what it does not:
  pinned (reference code is executed): experiments/scenarios.py:6-63 (observations), :124-192 (make_env: the
      flags collaborative=False, force_discrete_action=True, discrete_action, make_world(num_agents=n));
      experiments/run.py (loops), rls/* (Trainer, ActorNetwork, ReplayBuffer).
  still restated (upstream arithmetic, not in the reference tree): World.step, _set_action, the collision force,
      integrate_state, Scenario.reward / reset_world.
"""
