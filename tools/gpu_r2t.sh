# tools/gpu_r2t.sh : round 2 -- valued masked dot: 32-bit key tables + position tables read on a hit
set -x
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py -m gpu -x -q -k "dot or tri or ktruss or masked or nan or golden" 2>&1 | tail -4 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 3 --only default,valued,valued_old --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-420 $O/ab_tri_s22.log | tail -5
