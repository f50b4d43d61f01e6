# tools/gpu_r2z.sh : round 2 -- the warp-per-vector hash kernel with three warps per block (Erdos-Renyi line), the masked saxpy C<L>=L*L beside the dot at scale 22
set -x
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 600 python bench.py --workload spgemm --steps 5 --no-e2e --no-api --no-secondary > $O/bench_er20.json 2> $O/bench_er20.err
GB200_SAXPY_HASH_WARP=0 timeout 600 python bench.py --workload spgemm --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_er20_blockkernel.json 2> $O/bench_er20_blockkernel.err
python tools/show_bench.py $O/bench_er20.json $O/bench_er20_blockkernel.json | cut -c1-200
timeout 300 tools/launches.sh $O/spgemm_er20_launches.csv --workload spgemm
timeout 600 python tools/ab_tri.py --scale 22 --reps 2 --only default,masked_saxpy --out $O/ab_tri_s22_saxpy.json > $O/ab_tri_s22_saxpy.log 2>&1
cut -c1-330 $O/ab_tri_s22_saxpy.log | tail -3
