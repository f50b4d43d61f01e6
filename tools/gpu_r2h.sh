# tools/gpu_r2h.sh : round 2 -- operand residency cache (coherence test, resident t_api) + default bench
set -x
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cache.py tests/test_gpu_fullsize.py tests/test_demo_programs.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest_cache.log
cat $O/pytest_cache.log
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "rc=$?"; tail -5 $O/bench_default.err; cut -c1-1500 $O/bench_default.json
python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-250
du -sh $O
