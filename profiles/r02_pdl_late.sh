#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_pdl_late.txt; : > $out
for mode in 0 2 1; do for B in 7500 15000 30000 60000; do
  echo -n "MLP_PDL=$mode " >> $out
  BLA_MLP_PDL=$mode timeout 120 python profiles/step_prof.py $B 300 >> $out 2>&1
done; done
cat $out
