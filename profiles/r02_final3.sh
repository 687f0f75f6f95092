#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "import json; d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], json.dumps(d['e2e'])[:420])"
