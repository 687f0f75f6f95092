#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_gpu.py tests/test_programs_gpu.py -m gpu -x -q -k "hinge" 2>&1 | tail -3
timeout 120 python profiles/hinge_time.py 2>&1 | tail -1
