#!/bin/bash
# round 2, GPU call 17 (1 GPU): the expanded 1-D kinds (thermal / gamma / bounce-back) on the device
mkdir -p gpurun_out; O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_1d.py -m gpu -x -q ) > $O/pytest_call17.log 2>&1; tail -30 $O/pytest_call17.log
