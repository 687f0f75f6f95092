#!/bin/bash
# 2 GPUs: multi-GPU parity tests (peer windows one round / two rounds / NCCL) and a short strong-scaling bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r02_multigpu_n2.log 2>&1
tail -5 gpurun_out/r02_multigpu_n2.log
BLA_PEER_TWO_ROUNDS=1 DP_PATH=fp32 DP_BATCH=2000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dp_check.py > gpurun_out/r02_dp_check_n2_two_rounds.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 --no-extras > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
BLA_PEER_TWO_ROUNDS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 200 --warmup 20 --no-extras > gpurun_out/r02_bench_n2_two_rounds.json 2> gpurun_out/r02_bench_n2_two_rounds.err
tail -c 600 gpurun_out/r02_bench_n2.json
