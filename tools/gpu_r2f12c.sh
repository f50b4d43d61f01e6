# tools/gpu_r2f12c.sh : round 2 -- the device transpose with the staged scatter and the pattern shortcut: its GPU tests, then the A/B line at scale 22
set -x
O=gpurun_out/r2f12c; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py -m gpu -q -k "transpose" --tb=short -p no:cacheprovider 2>&1 | tail -4 > $O/pytest_transpose.log; cat $O/pytest_transpose.log
timeout 100 python tools/transpose_bench.py --scale 22 --check-scale 14 --ab --out $O/transpose_s22_ab.json > $O/transpose_s22_ab.log 2>&1; echo "rc=$?"; tail -1 $O/transpose_s22_ab.log | cut -c1-1500
