set -x
mkdir -p gpurun_out/r1a
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1a/pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1a/bench_tri_s22.json 2> gpurun_out/r1a/bench_tri_s22.err
python bench.py --workload sssp --steps 5 --warmup 3 --no-cpu > gpurun_out/r1a/bench_sssp_s22.json 2> gpurun_out/r1a/bench_sssp.err
python bench.py --workload bfs --steps 3 --warmup 3 --no-cpu > gpurun_out/r1a/bench_bfs_s22.json 2> gpurun_out/r1a/bench_bfs.err
python bench.py --workload bfs --bfs-dir pull --steps 3 --warmup 3 --no-cpu > gpurun_out/r1a/bench_bfspull_s22.json 2> gpurun_out/r1a/bench_bfspull.err
python bench.py --workload spgemm --steps 3 --warmup 3 --no-cpu > gpurun_out/r1a/bench_spgemm_er20.json 2> gpurun_out/r1a/bench_spgemm.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --warmup 3 --no-cpu > gpurun_out/r1a/bench_spgemm_rmat16.json 2> gpurun_out/r1a/bench_spgemm_rmat16.err
python bench.py --workload spgemm_rmat --scale 18 --steps 2 --warmup 2 --no-cpu > gpurun_out/r1a/bench_spgemm_rmat18.json 2> gpurun_out/r1a/bench_spgemm_rmat18.err
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1a/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1a/launches_tri_s22.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1a/ncu.log 2>&1
tail -3 gpurun_out/r1a/*.err gpurun_out/r1a/pytest.log
