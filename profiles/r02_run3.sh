#!/bin/bash
# one B200: full GPU suite, U-Net timing with programmatic launches everywhere, launch list of a 7,500-column shard step,
# ncu --set full of the dominant GEMM launch (256 x 784 x 56,832, integer pixels) and of group norm backward
mkdir -p gpurun_out
out=gpurun_out/r02_run3.txt; : > $out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1
tail -3 gpurun_out/r02_gputest.log >> $out
UNET_TIME=1 timeout 200 python profiles/unet_prof.py 64 3 tc >> $out 2>&1
UNET_TIME=1 timeout 200 python profiles/unet_prof.py 256 3 tc >> $out 2>&1
timeout 120 python profiles/gn_time.py >> $out 2>&1
for B in 7500 60000; do timeout 120 python profiles/step_prof.py $B 300 >> $out 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mlp_step_7500cols.csv \
    python profiles/step_prof.py 7500 6 > gpurun_out/ncu_7500.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_3xtf32 -s 2 -c 1 -f -o gpurun_out/r02_ncu_gemm_fwd1_main \
    python profiles/prof_one.py 0 0 256 56832 784 3xtf32 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:group_norm_bwd -s 1 -c 1 -f -o gpurun_out/r02_ncu_group_norm_bwd \
    python profiles/gn_time.py > gpurun_out/ncu_gn_bwd.log 2>&1
