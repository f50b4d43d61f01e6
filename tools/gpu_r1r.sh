set -x
O=gpurun_out/r1r; mkdir -p $O
GB200_TRACE=1 python bench.py --steps 5 --no-cpu > $O/bench_tri22.json 2> $O/bench_tri22.err
GB200_TRACE=1 python bench.py --steps 5 --no-cpu > $O/bench_tri22_b.json 2> $O/bench_tri22_b.err
grep -c . $O/*.err; tail -n 12 $O/bench_tri22.err $O/bench_tri22_b.err
