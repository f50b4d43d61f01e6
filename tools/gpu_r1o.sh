set -x
O=gpurun_out/r1o; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
python bench.py --steps 3 --no-cpu > $O/bench_tri22.json 2> $O/bench_tri22.err
python bench.py --workload tri --scale 20 --steps 3 --no-cpu --no-e2e > $O/bench_tri20.json 2> $O/bench_tri20.err
python bench.py --workload sssp --steps 5 --no-cpu --no-e2e > $O/bench_sssp.json 2> $O/bench_sssp.err
GB200_SPMV_OCC8=1 python bench.py --workload sssp --steps 5 --no-cpu --no-e2e > $O/bench_sssp_occ8.json 2> $O/bench_sssp_occ8.err
GB200_SPMV_GRID=12 python bench.py --workload sssp --steps 5 --no-cpu --no-e2e > $O/bench_sssp_g12.json 2> $O/bench_sssp_g12.err
python bench.py --workload bfs --steps 3 --no-cpu --no-e2e > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
python bench.py --workload spgemm_rmat --scale 18 --ef 8 --steps 2 --no-cpu --no-e2e > $O/bench_spgemm18.json 2> $O/bench_spgemm18.err
tools/prof.sh $O tri22 dotg_kernel 4 --workload tri --scale 22
tail -n 5 $O/*.err; cat $O/pytest.log
