# tools/gpu_r2d.sh : round 2 -- row-walk kernels v2 (branch-free probes, 4-row unroll) + two-pass set-up
set -x
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py -m gpu -x -q -k "dot or tri or golden" 2>&1 | tail -8 > $O/pytest_dot.log
cat $O/pytest_dot.log
timeout 400 python tools/ab_tri.py --scale 22 --reps 3 --only default,old,valued,valued_old,chunk512,chunk2048,hub4096,hub16384 --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -12
timeout 600 tools/prof.sh $O tri_s22 'dotr_kernel|dotg_kernel|dot_kernel' 5 --workload tri --scale 22
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
rm -f $O/plain_*.log $O/ncu_*.log
du -sh $O
