set -x
O=gpurun_out/r1m; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu > $O/bench_tri22_n2.json 2> $O/bench_tri22_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --workload sssp --no-cpu > $O/bench_sssp_n2.json 2> $O/bench_sssp_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 --workload spgemm_rmat --scale 16 --no-cpu > $O/bench_spgemm16_n2.json 2> $O/bench_spgemm16_n2.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 --no-cpu --no-e2e > $O/bench_spgemm16.json 2> $O/bench_spgemm16.err
python bench.py --workload spgemm_rmat --scale 18 --ef 8 --steps 2 --no-cpu --no-e2e > $O/bench_spgemm18.json 2> $O/bench_spgemm18.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
tail -n 5 $O/*.err; cat $O/pytest.log
