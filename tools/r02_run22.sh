#!/bin/bash
# round 2, GPU call 22 (1 GPU): the slab runtime's host loop on one rank
mkdir -p gpurun_out; O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_call22.log 2>&1; tail -15 $O/pytest_call22.log
