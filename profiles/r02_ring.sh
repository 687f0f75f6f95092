#!/bin/bash
# ring geometry by tile width (deeper rings for narrow tiles) against the fixed full-width geometry
mkdir -p gpurun_out
out=gpurun_out/r02_ring.txt; : > $out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_ring.log 2>&1
tail -3 gpurun_out/r02_gputest_ring.log >> $out
for fx in 1 0; do
  for B in 7500 15000 30000 60000; do
    echo -n "RING_FIXED=$fx " >> $out
    BLA_TC_RING_FIXED=$fx timeout 120 python profiles/step_prof.py $B 300 >> $out 2>&1
  done
  echo -n "RING_FIXED=$fx " >> $out
  BLA_TC_RING_FIXED=$fx UNET_TIME=1 timeout 200 python profiles/unet_prof.py 64 3 tc 2>&1 | grep "train step" >> $out
done
cat $out
