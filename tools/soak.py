"""Soak run at the north-star size: 1,048,576 simple_spread envs, 10,000 fused steps (400 episodes per env) in
25-step rollout calls on one B200.  Checks that the episode statistics stay sane (every env finishes every episode, no
non-finite return beyond the coincident-agent cases upstream's 0/0 produces, mean return stationary) and reports the
sustained throughput with the clocks and power sampled over the whole run.  Prints one JSON object."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiagent_rl_b200 as m  # noqa: E402
from multiagent_rl_b200.networks import random_state_dict  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Clocks  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
CALLS = int(sys.argv[2]) if len(sys.argv) > 2 else 400
env = m.make_env('simple_spread', num_envs=B, batched=True, seed=12345678, max_episode_len=25)
actor = m.FusedActor(random_state_dict(10, 5, 12345678), seed=12345678)
env.reset()
env.rollout(actor, 25)
torch.cuda.synchronize()
env.read_stats(clear=True)
clk = Clocks(0, period_s=0.05)
clk.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
means = []
for c in range(CALLS):
    env.rollout(actor, 25)
    if (c + 1) % 100 == 0:
        s = env.read_stats(clear=True)
        means.append(s[0] / max(s[2] - s[4], 1))
        assert s[2] == 100.0 * B and s[3] == 2500.0 * B, s
        assert s[4] <= 1e-5 * s[2], ('non-finite episodes', s[4])
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) * 1e-3
out = {'envs': B, 'steps': CALLS * 25, 'seconds': dt, 'wall_seconds': time.perf_counter() - t0,
       'G_agent_steps_per_s': B * 3 * CALLS * 25 / dt / 1e9, 'mean_return_per_100_episodes': means,
       'clocks': clk.stop(), 'finite_state': bool(torch.isfinite(env.observe()).all())}
assert max(means) - min(means) < 0.5, means
print(json.dumps(out))
