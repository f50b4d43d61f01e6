set -x
O=gpurun_out/r1g; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
tools/prof.sh $O tri20 dotg_kernel 2 --workload tri --scale 20
python bench.py --workload tri --scale 20 --steps 3 --no-cpu > $O/bench_tri20.json 2> $O/bench_tri20.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_tri20.csv python bench.py --workload tri --scale 20 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncul_tri20.log 2>&1
tools/prof.sh $O spgemm16 'saxpy_light|saxpy_heavy|sym_hash|heavy_' 24 --workload spgemm_rmat --scale 16
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_spgemm16.csv python bench.py --workload spgemm_rmat --scale 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/ncul_spgemm16.log 2>&1
tools/prof.sh $O sssp spmv_stream 1 --workload sssp
tools/prof.sh $O bfs 'saxpyv' 6 --workload bfs
du -sh $O; tail -n 3 $O/*.err $O/pytest.log
