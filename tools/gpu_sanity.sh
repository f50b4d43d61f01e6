# tools/gpu_sanity.sh TAG : what the driver runs at round end, in short
set -x
TAG=${1:-sanity}
O=gpurun_out/$TAG; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
python bench.py --workload bfs --steps 3 --no-cpu > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
cat $O/pytest.log; tail -2 $O/smoke.log; tail -n 3 $O/*.err
