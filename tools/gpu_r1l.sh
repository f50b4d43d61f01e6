set -x
O=gpurun_out/r1l; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
GB200_HEAVY_L2_MB=100000 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16_nol2.json 2> $O/bench_spgemm16_nol2.err
GB200_HEAVY_L2_MB=40 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16_l2_40.json 2> $O/bench_spgemm16_l2_40.err
GB200_HEAVY_L2_MB=160 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16_l2_160.json 2> $O/bench_spgemm16_l2_160.err
python bench.py --workload spgemm_rmat --scale 18 --ef 8 --steps 2 --no-cpu --no-e2e > $O/bench_spgemm18.json 2> $O/bench_spgemm18.err
tools/launches.sh $O/launches_spgemm16.csv --workload spgemm_rmat --scale 16
tools/prof.sh $O spgemm16 'saxpy_heavy|saxpy_light' 6 --workload spgemm_rmat --scale 16
du -sh $O; tail -n 3 $O/*.err; cat $O/pytest.log
