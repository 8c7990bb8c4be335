"""CPU oracle for the MPE step + observe + reward + actor-forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multiagent_rl_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU baseline being timed.

Parity status (see DESIGN.md "Oracle"):
  * physics / reward (``mpe_ref``, ``mpe_vec``): **parity unpinned**.  The
    arithmetic lives in the third-party ``multiagent`` package (OpenAI
    multi-agent-particle-envs lineage, imported at
    /root/reference/experiments/scenarios.py:2-3) which is not vendored, not
    pinned and not installable offline; the reference ships no tests or golden
    vectors.  The restatement follows the published upstream algorithm
    function by function and is anchored on hand-derived known answers.
  * observations (``local_obs_*``) and ``make_env``: follow /root/reference/experiments/scenarios.py:6-63,124-192
    and are **pinned by execution**: ``python -m oracle.build_ref`` byte-compiles the reference's modules into
    ``oracle/_ref`` and tests/test_reference_exec.py runs the reference's own ``make_env`` against this package
    (bit-exact on all fixtures).
  * MAAC-fork engine + ``fullobs_collect_treasure`` (``maac_ref``): **parity unpinned** for the fork-side arithmetic
    (shariqiqbal2810/multiagent-particle-envs, not in the reference tree); its observation restates
    /root/reference/experiments/scenarios.py:95-121 and is **pinned by execution** (tests/test_treasure_oracle.py).
  * actor forward (``actor_ref``): **pinned** against the reference's own
    ``rls.model.ac_network_multi_gumbel.ActorNetwork`` run in the authoring
    container (``oracle/gen_golden.py`` -> ``tests/golden/actor_*.npz``); critic forward (``critic_ref``) likewise
    (``tests/golden/critic_*.npz``).
"""
