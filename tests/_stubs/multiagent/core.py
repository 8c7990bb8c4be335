"""``multiagent.core`` - engine classes of the oracle's restatement (oracle/mpe_ref.py)."""
from oracle.mpe_ref import Action, Agent, AgentState, Entity, EntityState, Landmark, World  # noqa: F401
