set -x
O=gpurun_out/r1h; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
python bench.py --workload tri --scale 20 --steps 3 --no-cpu > $O/bench_tri20.json 2> $O/bench_tri20.err
GB200_DOTG_ISO=0 python bench.py --workload tri --scale 20 --steps 3 --no-cpu --no-e2e > $O/bench_tri20_noiso.json 2> $O/bench_tri20_noiso.err
python bench.py --steps 3 > $O/bench_tri22.json 2> $O/bench_tri22.err
GB200_DOTG_ISO=0 python bench.py --steps 3 --no-cpu --no-e2e > $O/bench_tri22_noiso.json 2> $O/bench_tri22_noiso.err
tools/prof.sh $O tri22 dotg_kernel 2 --workload tri --scale 22
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_tri22.csv python bench.py --workload tri --scale 22 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncul_tri22.log 2>&1
du -sh $O; tail -n 3 $O/*.err; cat $O/pytest.log
