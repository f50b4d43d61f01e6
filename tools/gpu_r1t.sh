set -x
O=gpurun_out/r1t; mkdir -p $O
GB200_TRACE=1 python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
GB200_TRACE=1 python bench.py --workload spgemm --steps 3 --no-cpu > $O/bench_spgemm_er20.json 2> $O/bench_spgemm_er20.err
tail -n 8 $O/*.err
