#!/bin/bash
# 2 GPUs: multi-GPU parity tests (peer windows one round / two rounds / NCCL) and a short strong-scaling bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r02_multigpu_n2.log 2>&1
tail -5 gpurun_out/r02_multigpu_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 --no-extras > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 600 gpurun_out/r02_bench_n2.json
