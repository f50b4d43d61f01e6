set -x
O=gpurun_out/r1n; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload sssp --no-cpu > $O/bench_sssp_n2.json 2> $O/bench_sssp_n2.err
python bench.py --workload sssp --steps 5 --no-cpu --no-e2e > $O/bench_sssp.json 2> $O/bench_sssp.err
python bench.py --workload bfs --steps 3 --no-cpu --no-e2e > $O/bench_bfs.json 2> $O/bench_bfs.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
python bench.py --workload spgemm_rmat --scale 18 --ef 8 --steps 2 --no-cpu --no-e2e > $O/bench_spgemm18.json 2> $O/bench_spgemm18.err
python bench.py --workload spgemm --steps 3 --no-cpu --no-e2e > $O/bench_spgemm_er20.json 2> $O/bench_spgemm_er20.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
tools/launches.sh $O/launches_spgemm18.csv --workload spgemm_rmat --scale 18 --ef 8
tools/launches.sh $O/launches_bfs.csv --workload bfs
tail -n 5 $O/*.err; cat $O/pytest.log
