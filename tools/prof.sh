#!/bin/bash
# tools/prof.sh OUTDIR NAME KERNEL_REGEX COUNT bench-args...
# One ncu --set full capture of the named kernel(s) of a bench.py run (after the same command has
# exited 0 without ncu), exported on the box as CSV (raw page + source page); the .ncu-rep itself
# stays on the box unless it is small (gpurun_out/ is capped at 64 MiB).
O=$1; name=$2; rx=$3; cnt=$4; shift 4
mkdir -p $O
python bench.py "$@" --steps 1 --warmup 1 --no-cpu --no-e2e --no-api --no-secondary > $O/plain_$name.log 2>&1 || { echo "plain run failed: $name"; tail -5 $O/plain_$name.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -f -o /tmp/prof_$name \
    python bench.py "$@" --steps 1 --warmup 1 --no-cpu --no-e2e --no-api --no-secondary > $O/ncu_$name.log 2>&1
ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page source --csv --print-source cuda,sass > $O/${name}_source.csv 2>/dev/null
sz=$(stat -c %s /tmp/prof_$name.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 12000000 ]; then cp /tmp/prof_$name.ncu-rep $O/; fi
# keep the source page bounded
if [ $(stat -c %s $O/${name}_source.csv) -gt 8000000 ]; then head -c 8000000 $O/${name}_source.csv > $O/${name}_source.csv.tmp; mv $O/${name}_source.csv.tmp $O/${name}_source.csv; fi
ls -la $O | grep $name
