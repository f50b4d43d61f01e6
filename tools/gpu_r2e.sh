# tools/gpu_r2e.sh : round 2 -- warp items for tiny owners
set -x
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py -m gpu -x -q -k "dot or tri or golden" 2>&1 | tail -8 > $O/pytest_dot.log
cat $O/pytest_dot.log
timeout 400 python tools/ab_tri.py --scale 22 --reps 3 --only default,notiny,old,valued,valued_old --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -8
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
du -sh $O
