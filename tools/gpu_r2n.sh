# tools/gpu_r2n.sh : round 2 -- side streams, L2 prefetch of the coming tasks, batched value loads, short-list trim skip
set -x
O=gpurun_out/r2n; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,memory.total --format=csv > $O/gpu.csv
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 3 --only default,r2m,nostreams,nopf,pf_hub_only,pf1,pf4,pf8,trim1,clsu1,valued,valued_r2m --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-420 $O/ab_tri_s22.log | tail -14
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
