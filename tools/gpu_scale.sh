# tools/gpu_scale.sh TAG : the multi-GPU lines (run under gpurun --gpus 8)
set -x
TAG=${1:-r1_scale}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
run () {  # n name args...
  n=$1; name=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu "$@" > $O/bench_${name}_n$n.json 2> $O/bench_${name}_n$n.err
}
run 8 tri_s22
run 4 tri_s22
run 2 tri_s22
python bench.py --steps 5 --no-cpu > $O/bench_tri_s22_n1.json 2> $O/bench_tri_s22_n1.err
run 8 spgemm_rmat16 --workload spgemm_rmat --scale 16
run 8 sssp_s22 --workload sssp
tail -n 4 $O/*.err
