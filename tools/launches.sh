#!/bin/bash
# tools/launches.sh OUT.csv bench-args... : per-launch durations of this library's kernels in one
# bench.py step (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: shares, not absolutes)
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gb200:: -c 3000 --csv \
    --log-file $out python bench.py "$@" --steps 1 --warmup 0 --no-cpu --no-e2e --no-api --no-secondary > ${out%.csv}.log 2>&1
