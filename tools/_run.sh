python -m pytest tests/test_gpu_actor.py -q -x -k "host_rollout_shards" 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0,'.')
import torch, multiagent_rl_b200 as m
for prec in ('fp32','fp64'):
    env = m.make_env('fullobs_collect_treasure', num_envs=77, batched=True, seed=3, precision=prec, max_episode_len=3)
    env.track_returns(True); env.reset()
    for t in range(5):
        env.step(torch.randint(0,5,(77,8),dtype=torch.int32,device='cuda'), info=True)
    env.reset(); env.observe(); env.read_stats()
torch.cuda.synchronize(); print('sanitize treasure done')
" 2>&1 | tail -6
