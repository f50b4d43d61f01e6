# tools/gpu_r2n2c.sh : round 2, two GPUs -- the BFS push step sharded over the GPUs, checked level by level
set -x
O=gpurun_out/r2n2c; mkdir -p $O
run () { name=$1; n=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" 2> $O/$name.err | grep '^{' | tail -1 > $O/$name.json; echo "rc=$? $name"; grep -v "^\*\|OMP_NUM\|^$" $O/$name.err | tail -4 | cut -c1-300; }
run bench_bfs_n2 2 --workload bfs --steps 10 --warmup 3 --no-cpu
python tools/show_bench.py $O/bench_*.json 2>/dev/null | cut -c1-220
grep -o '"exchange_check": "[^"]*"' $O/bench_bfs_n2.json
