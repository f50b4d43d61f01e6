# tools/gpu_r2j.sh : round 2 -- whole GPU suite (NaN semantics, GxB_select) + A/B of 8 rows per iteration in the hub kernel
set -x
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 400 python tools/ab_tri.py --scale 22 --reps 3 --only default,hub4096,valued --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -5
GB200_SHIM_STATS=1 timeout 300 python -m pytest tests/test_demo_programs.py -m gpu -x -q 2>&1 | tail -3
