# tools/gpu_r2trace.sh : round 2 -- the default bench line with GB200_TRACE=1: trips of the block cache to the driver inside the timed multiplies
set -x
O=gpurun_out/r2trace; mkdir -p $O
GB200_TRACE=1 timeout 600 python bench.py --no-cpu --no-e2e --no-api > $O/bench_default_trace.json 2> $O/bench_default_trace.err
grep "gb200_AxB_device" $O/bench_default_trace.err | awk '{print NR": "$0}' | cut -c1-170 | sed -n '1,12p;13,40p'
python tools/show_bench.py $O/bench_default_trace.json | cut -c1-200
