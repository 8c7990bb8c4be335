#!/bin/bash
# Round-2 ncu evidence, part 2: the large-team kernels (actor k_tc2<0,6|12,0,8>, k_step_grp, fused large-team rollout)
# and the critic.  One gpurun call.
set -x
cd "$(dirname "$0")/.."
python tools/profile_target.py bign > gpurun_out/r2_plain_bign.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_step_grp|k_tc2|k_reset_grp" -c 12 -o gpurun_out/r2_prof_bign python tools/profile_target.py bign > /dev/null 2>&1
python tools/profile_target.py critic > gpurun_out/r2_plain_critic.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_critic|k_dense3" -c 4 -o gpurun_out/r2_prof_critic python tools/profile_target.py critic > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
