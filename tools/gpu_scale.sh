# tools/gpu_scale.sh TAG : the multi-GPU lines of the default workload (run under gpurun --gpus 8)
set -x
TAG=${1:-r1_scale}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
run () {  # n name args...
  n=$1; name=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu "$@" > $O/bench_${name}_n$n.json 2> $O/bench_${name}_n$n.err
}
run 8 tri_s22
run 4 tri_s22
tail -n 4 $O/*.err
