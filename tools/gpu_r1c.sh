mkdir -p gpurun_out/r1c
python -m pytest tests -m gpu -x -q -k "vector or bfs or sssp or mxv" 2>&1 | tail -5 > gpurun_out/r1c/pytest.log
for g in 4 8 16 32; do
GB200_DOTV_G=$g python bench.py --workload sssp --steps 5 --warmup 3 --no-cpu > gpurun_out/r1c/bench_sssp_s22_G$g.json 2> gpurun_out/r1c/bench_sssp.err
done
python bench.py --workload bfs --steps 3 --warmup 3 --no-cpu > gpurun_out/r1c/bench_bfs_s22.json 2> gpurun_out/r1c/bench_bfs.err
python bench.py --workload bfs --bfs-dir pull --steps 3 --warmup 3 --no-cpu > gpurun_out/r1c/bench_bfspull_s22.json 2> gpurun_out/r1c/bench_bfspull.err
cat gpurun_out/r1c/pytest.log; tail -n 3 gpurun_out/r1c/*.err
