#!/bin/bash
# round 2, GPU call 13 (1 GPU): device timeline of swalbe_time_loop_host (SWALBE_HOST_TRACE), copies alone, loop alone
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python tools/e2e_probe.py --steps 20 --bands 0,2048 --kmax 0 --trace > $O/probes13.txt 2>&1; cat $O/probes13.txt
