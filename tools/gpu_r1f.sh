set -x
O=gpurun_out/r1f; mkdir -p $O
python bench.py --steps 3 --no-cpu > $O/bench_tri.json 2> $O/bench_tri.err
prof () {  # name, kernel regex, count, bench args...
  name=$1; rx=$2; cnt=$3; shift 3
  python bench.py "$@" --steps 1 --warmup 1 --no-cpu > $O/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -o $O/prof_$name python bench.py "$@" --steps 1 --warmup 1 --no-cpu > $O/ncu_$name.log 2>&1
}
prof sssp spmv_stream 2 --workload sssp
prof bfs 'saxpyv' 12 --workload bfs
prof spgemm16 'saxpy_light|saxpy_heavy|sym_hash|heavy_' 24 --workload spgemm_rmat --scale 16
prof tri20 dotg_kernel 2 --workload tri --scale 20
python bench.py --workload tri --scale 20 --steps 1 --warmup 1 --no-cpu > $O/plain_tri20b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_tri20.csv python bench.py --workload tri --scale 20 --steps 1 --warmup 1 --no-cpu > $O/ncul_tri20.log 2>&1
tail -n 3 $O/*.err $O/ncu_*.log
