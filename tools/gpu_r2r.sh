# tools/gpu_r2r.sh : round 2 -- the whole GPU suite and the default bench line with the final code
set -x
O=gpurun_out/r2r; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/gpu.csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
echo "rc=$?"; tail -3 $O/bench_default.err; cat $O/bench_default.time; python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-220
( time timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/bench_reference.time
echo "rc=$?"; cat $O/bench_reference.time; cut -c1-600 $O/bench_reference.json
