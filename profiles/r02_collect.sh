#!/bin/bash
# Round-2 evidence collection on one B200 (run under gpurun): bench line, launch lists, ncu captures.
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
UNET_TIME=1 python profiles/unet_prof.py 64 3 tc > gpurun_out/r02_unet_time.txt 2>&1
python profiles/gn_time.py > gpurun_out/r02_gn_time.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mlp_step.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_mlp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_unet_step.csv \
    python profiles/unet_prof.py 64 2 tc > gpurun_out/ncu_unet.log 2>&1
for w in fwd bwd; do
ncu --set full --clock-control none --import-source on -k regex:group_norm_${w} -s 1 -c 1 -f -o gpurun_out/r02_ncu_group_norm_${w} \
    python profiles/gn_time.py > gpurun_out/ncu_gn_${w}.log 2>&1
done
