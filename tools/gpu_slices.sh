# tools/gpu_slices.sh TAG K : every rank's share of a K-way run of the default workload, one after the other on one GPU
set -x
TAG=${1:-slices}; K=${2:-8}
O=gpurun_out/$TAG; mkdir -p $O
for r in $(seq 0 $((K-1))); do python bench.py --steps 5 --no-cpu --no-e2e --slice-of $K --slice-rank $r > $O/bench_tri22_slice$r.json 2> $O/bench_tri22_slice$r.err; done
tail -n 3 $O/*.err
