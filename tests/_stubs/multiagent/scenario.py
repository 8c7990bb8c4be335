"""``multiagent.scenario`` - upstream's BaseScenario."""
from oracle.mpe_ref import BaseScenario  # noqa: F401
