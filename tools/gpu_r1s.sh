set -x
O=gpurun_out/r1s; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
GB200_TRACE=1 python bench.py --steps 5 --no-cpu > $O/bench_tri22.json 2> $O/bench_tri22.err
python bench.py --workload tri --scale 20 --steps 5 --no-cpu > $O/bench_tri20.json 2> $O/bench_tri20.err
GB200_DOTG_ISO=0 python bench.py --steps 3 --no-cpu --no-e2e > $O/bench_tri22_valued.json 2> $O/bench_tri22_valued.err
python bench.py --workload sssp --steps 5 --no-cpu > $O/bench_sssp.json 2> $O/bench_sssp.err
python bench.py --workload bfs --steps 3 --no-cpu > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
tools/prof.sh $O tri22 dotg_kernel 4 --workload tri --scale 22
tail -n 8 $O/*.err; cat $O/pytest.log
