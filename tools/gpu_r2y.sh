# tools/gpu_r2y.sh : round 2 -- hub walk with software-pipelined groups of four rows (A/B against eight rows per iteration)
set -x
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py -m gpu -x -q -k "masked_dot or golden or tri_demo" 2>&1 | tail -3 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 4 --only default,nopipe --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -4
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
