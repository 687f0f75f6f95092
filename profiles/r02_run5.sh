#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run5.txt; : > $out
for occ in 3 4; do for B in 15000 30000 60000; do
  echo -n "HEAD_OCC=$occ " >> $out
  BLA_HEAD_OCC=$occ timeout 120 python profiles/step_prof.py $B 300 >> $out 2>&1
done; done
