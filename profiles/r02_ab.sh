#!/bin/bash
# A/B of the round-2 launch-structure changes on one B200: PDL and the forked layer-2 weight gradient, shard sizes of 8/4/2/1 GPUs.
mkdir -p gpurun_out
out=gpurun_out/r02_ab.txt; : > $out
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  for B in 7500 15000 30000 60000; do
    echo -n "PDL=$1 FORK=$2 " >> $out
    BLA_PDL=$1 BLA_MLP_FORK=$2 timeout 120 python profiles/step_prof.py $B 300 >> $out 2>&1
  done
done
timeout 120 python profiles/gn_time.py >> $out 2>&1
UNET_TIME=1 timeout 200 python profiles/unet_prof.py 64 3 tc >> $out 2>&1
BLA_PDL=0 UNET_TIME=1 timeout 200 python profiles/unet_prof.py 64 3 tc >> $out 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1
tail -3 gpurun_out/r02_gputest.log >> $out
