# tools/gpu_r2l.sh : round 2 -- slab-streamed unmasked C=A*A (checksum and discard): RMAT 18 in forced slabs
# against the single-call line, then RMAT 20 on one GPU; cache / reduce tests
set -x
O=gpurun_out/r2l; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cache.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload spgemm_rmat --scale 18 --ef 16 --slab-gb 10 --steps 2 --warmup 1 --no-secondary > $O/bench_rmat18_slabs.json 2> $O/bench_rmat18_slabs.err
echo "rc=$?"; tail -4 $O/bench_rmat18_slabs.err | cut -c1-300
timeout 900 python bench.py --workload spgemm_rmat --scale 18 --ef 16 --steps 2 --warmup 1 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_rmat18_whole.json 2> $O/bench_rmat18_whole.err
timeout 1500 python bench.py --workload spgemm_rmat --scale 20 --ef 16 --steps 2 --warmup 1 --no-secondary > $O/bench_rmat20_slabs.json 2> $O/bench_rmat20_slabs.err
echo "rc=$?"; tail -4 $O/bench_rmat20_slabs.err | cut -c1-300
python tools/show_bench.py $O/bench_*.json 2>/dev/null | cut -c1-220
