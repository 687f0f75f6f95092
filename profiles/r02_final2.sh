#!/bin/bash
# last collection of the round on one B200: the default bench line and the ncu launch list of the same step
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mlp_step.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_mlp.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
