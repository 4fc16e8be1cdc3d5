#!/bin/bash
# GPU session 18 (round 2, 2 GPUs): the C host program end to end on a synthetic channel, 1 and 2 GPUs.
mkdir -p gpurun_out /tmp/cli && cd /tmp/cli
python $GRAFT_REPO_ROOT/tools/make_inputs.py channel /tmp/cli --nx 8192 --ny 4096 --iters 401 > $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log 2>&1
ls /tmp/cli >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
P=$(ls /tmp/cli/*.params | head -1); O=$(ls /tmp/cli/obstacles* | head -1)
for g in 1 2; do
  echo "== LBM_GPUS=$g" >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
  LBM_GPUS=$g LBM_SKIP_FINAL_STATE=1 LBM_REPORT=1 timeout 300 $GRAFT_REPO_ROOT/d2q9-bgk $P $O >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log 2>&1
  echo "rc=$?" >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
  md5sum av_vels.dat >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
done
echo "== LBM_KERNEL=vec4 (one-step kernel), 1 GPU" >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
LBM_KERNEL=vec4 LBM_SKIP_FINAL_STATE=1 LBM_REPORT=1 timeout 300 $GRAFT_REPO_ROOT/d2q9-bgk $P $O >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log 2>&1
md5sum av_vels.dat >> $GRAFT_REPO_ROOT/gpurun_out/s18_cli.log
echo done
