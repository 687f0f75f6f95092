#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_pack.log 2>&1; tail -3 gpurun_out/r02_gputest_pack.log
timeout 300 python bench.py --no-extras --no-cpu --steps 100 > gpurun_out/r02_bench_pack.json 2> gpurun_out/r02_bench_pack.err
python -c "import json; d=json.loads(open('gpurun_out/r02_bench_pack.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], json.dumps(d['e2e'])[:700]); print(d['e2e_u8']['ms_per_step'])"
nproc
