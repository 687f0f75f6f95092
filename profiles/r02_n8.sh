#!/bin/bash
# 8 GPUs: data-parallel parity (dp_check: default all-reduce = two-round peer kernel from 4 ranks), then the strong-scaling bench line with
# the two-round and the one-round peer kernels
mkdir -p gpurun_out
DP_PATH=fp32 DP_BATCH=4096 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/dp_check.py > gpurun_out/r02_dp_check_n8.log 2>&1
grep -E "OK|peer_windows|FAIL|Error" gpurun_out/r02_dp_check_n8.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 200 --warmup 20 --no-extras > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -c 700 gpurun_out/r02_bench_n8.json
