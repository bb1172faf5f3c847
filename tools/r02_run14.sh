#!/bin/bash
# round 2, GPU call 14 (1 GPU): cost of band launches (no copies) for forced geometries
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python tools/band_probe.py > $O/probes14.txt 2>&1; cat $O/probes14.txt
