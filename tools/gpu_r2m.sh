# tools/gpu_r2m.sh : round 2 -- GPU suite + tri A/B after the one-pass item lists
set -x
O=gpurun_out/r2m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 400 python tools/ab_tri.py --scale 22 --reps 3 --only default,old,valued --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -4
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
