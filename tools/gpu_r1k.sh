set -x
O=gpurun_out/r1k; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
python bench.py --workload tri --scale 20 --steps 3 --no-cpu > $O/bench_tri20.json 2> $O/bench_tri20.err
python bench.py --steps 3 --no-cpu > $O/bench_tri22.json 2> $O/bench_tri22.err
GB200_DOTG_ISO=0 python bench.py --steps 3 --no-cpu --no-e2e > $O/bench_tri22_noiso.json 2> $O/bench_tri22_noiso.err
python bench.py --workload sssp --steps 5 --no-cpu > $O/bench_sssp.json 2> $O/bench_sssp.err
python bench.py --workload bfs --steps 3 --no-cpu > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
GB200_HEAVY_L2_MB=100000 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16_nol2.json 2> $O/bench_spgemm16_nol2.err
GB200_SYM_BITMAP=0 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16_nobm.json 2> $O/bench_spgemm16_nobm.err
python bench.py --workload spgemm --steps 3 --no-cpu > $O/bench_spgemm_er20.json 2> $O/bench_spgemm_er20.err
tools/launches.sh $O/launches_tri22.csv --workload tri --scale 22
tools/launches.sh $O/launches_spgemm16.csv --workload spgemm_rmat --scale 16
du -sh $O; tail -n 3 $O/*.err; cat $O/pytest.log
