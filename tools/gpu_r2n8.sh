# tools/gpu_r2n8.sh : round 2, eight GPUs -- tri and SSSP scaling lines, the peer exchange check at N = 8
set -x
O=gpurun_out/r2n8; mkdir -p $O
run () { name=$1; n=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" 2> $O/$name.err | grep '^{' | tail -1 > $O/$name.json; echo "rc=$? $name"; tail -2 $O/$name.err | cut -c1-300; }
run bench_tri_n8 8 --steps 5 --warmup 3 --no-cpu
run bench_sssp_n8 8 --workload sssp --steps 10 --warmup 3 --no-cpu
run bench_tri_n4 4 --steps 5 --warmup 3 --no-cpu
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/peer_check.py --scale 18 --rounds 3 > $O/peer_check_n8.log 2>&1
tail -5 $O/peer_check_n8.log
python tools/show_bench.py $O/bench_*.json 2>/dev/null | cut -c1-220
