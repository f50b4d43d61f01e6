set -x
O=gpurun_out/r1n2; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > $O/bench_tri22_n2.json 2> $O/bench_tri22_n2.err
tail -n 5 $O/*.err
