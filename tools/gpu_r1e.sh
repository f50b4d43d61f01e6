set -x
O=gpurun_out/r1e; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.log
python bench.py > $O/bench_tri.json 2> $O/bench_tri.err
python bench.py --workload sssp > $O/bench_sssp.json 2> $O/bench_sssp.err
GB200_SPMV_STREAM=0 python bench.py --workload sssp --no-cpu > $O/bench_sssp_old.json 2> $O/bench_sssp_old.err
python bench.py --workload bfs --steps 3 > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload bfs --bfs-dir pull --steps 3 --no-cpu > $O/bench_bfspull.json 2> $O/bench_bfspull.err
python bench.py --workload spgemm --steps 3 > $O/bench_spgemm_er20.json 2> $O/bench_spgemm_er20.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 > $O/bench_spgemm_rmat16.json 2> $O/bench_spgemm_rmat16.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -n 3 $O/*.err $O/pytest.log
