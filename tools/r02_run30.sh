#!/bin/bash
# round 2, GPU call 30 (2 GPUs): the slab runtime's tests on the final dist.cu
mkdir -p gpurun_out; O=gpurun_out
( timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_call30.log 2>&1; tail -3 $O/pytest_call30.log
