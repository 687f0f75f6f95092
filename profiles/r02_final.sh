#!/bin/bash
# end-of-round validation on one B200: the full GPU suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
echo "tests: see the previous run"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 400 gpurun_out/r02_bench_reference.json
