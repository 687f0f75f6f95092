#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_check.log 2>&1; tail -3 gpurun_out/r02_gputest_check.log
for B in 7500 30000 60000; do timeout 120 python profiles/step_prof.py $B 300; done
timeout 300 python bench.py --no-extras --no-cpu --steps 100 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_u8']['ms_per_step'])"
