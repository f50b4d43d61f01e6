mkdir -p gpurun_out/r1b
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b/pytest.log
python bench.py --workload sssp --steps 5 --warmup 3 --no-cpu > gpurun_out/r1b/bench_sssp_s22.json 2> gpurun_out/r1b/bench_sssp.err
python bench.py --workload bfs --steps 3 --warmup 3 --no-cpu > gpurun_out/r1b/bench_bfs_s22.json 2> gpurun_out/r1b/bench_bfs.err
python bench.py --workload bfs --bfs-dir pull --steps 3 --warmup 3 --no-cpu > gpurun_out/r1b/bench_bfspull_s22.json 2> gpurun_out/r1b/bench_bfspull.err
cat gpurun_out/r1b/pytest.log; tail -n 3 gpurun_out/r1b/*.err
