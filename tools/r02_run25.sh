#!/bin/bash
# round 2, GPU call 25 (1 GPU): compute-sanitizer over what the second half of the round added
mkdir -p gpurun_out; O=gpurun_out
( time timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_r02b.py ) > $O/sanitize_r02b_memcheck.log 2>&1; tail -6 $O/sanitize_r02b_memcheck.log
( time timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_r02b.py ) > $O/sanitize_r02b_racecheck.log 2>&1; tail -6 $O/sanitize_r02b_racecheck.log
