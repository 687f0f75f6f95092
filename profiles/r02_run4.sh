#!/bin/bash
# one B200: full GPU suite; U-Net step with skinny products on the tensor path; ncu captures of the hinge pass and the MLP head kernel
mkdir -p gpurun_out
out=gpurun_out/r02_run4.txt; : > $out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1
tail -3 gpurun_out/r02_gputest.log >> $out
for t in 1e8 2e7 5e6; do
  echo "BLA_TC_SKINNY_MIN=$t" >> $out
  BLA_TC_SKINNY_MIN=$t UNET_TIME=1 timeout 200 python profiles/unet_prof.py 64 3 tc 2>&1 | grep "train step" >> $out
done
ncu --set full --clock-control none --import-source on -k regex:hinge_onepass -s 1 -c 1 -f -o gpurun_out/r02_ncu_hinge_onepass \
    python profiles/hinge_prof.py > gpurun_out/ncu_hinge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_train -s 3 -c 1 -f -o gpurun_out/r02_ncu_head_train \
    python profiles/step_prof.py 60000 4 > gpurun_out/ncu_head.log 2>&1
