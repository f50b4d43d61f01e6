set -x
O=gpurun_out/r1q; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
python bench.py --steps 5 --no-cpu > $O/bench_tri22.json 2> $O/bench_tri22.err
python bench.py --workload tri --scale 20 --steps 5 --no-cpu > $O/bench_tri20.json 2> $O/bench_tri20.err
python bench.py --workload sssp --steps 5 --no-cpu --no-e2e > $O/bench_sssp.json 2> $O/bench_sssp.err
tools/launches.sh $O/launches_tri22.csv --workload tri --scale 22
tail -n 5 $O/*.err; cat $O/pytest.log
