# tools/gpu_r2n8b.sh : round 2, eight GPUs -- tri / SSSP / BFS lines with the final kernels, and the
# flop-balanced B-column split of C=A*A (slab-streamed where C exceeds HBM) at RMAT scale 18, 20 and 22
set -x
O=gpurun_out/r2n8b; mkdir -p $O
run () { name=$1; n=$2; to=$3; shift 3; timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" 2> $O/$name.err | grep '^{' | tail -1 > $O/$name.json; echo "rc=$? $name"; grep -v "^\*\|OMP_NUM\|^$" $O/$name.err | tail -3 | cut -c1-300; }
run bench_tri_n8 8 150 --steps 5 --warmup 3 --no-cpu
run bench_sssp_n8 8 120 --workload sssp --steps 10 --warmup 3 --no-cpu
run bench_bfs_n8 8 120 --workload bfs --steps 10 --warmup 3 --no-cpu
run bench_spgemm_rmat20_n8 8 150 --workload spgemm_rmat --scale 20 --steps 2 --warmup 1 --no-cpu
run bench_spgemm_rmat22_n8 8 180 --workload spgemm_rmat --scale 22 --steps 1 --warmup 1 --no-cpu
python tools/show_bench.py $O/bench_*.json 2>/dev/null | cut -c1-220
