# tools/gpu_r2f12a.sh : round 2, rows f1 / f2 on the GPU -- the whole GPU suite (new tests included), smoke(), the device transpose at scale 22
set -x
O=gpurun_out/r2f12; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=8 --tb=short -p no:cacheprovider > $O/pytest_gpu_full.log 2>&1 ) 2> $O/pytest_gpu.time
tail -4 $O/pytest_gpu_full.log; grep -E "^(FAILED|ERROR)" $O/pytest_gpu_full.log | head -12; grep real $O/pytest_gpu.time
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 200 python tools/transpose_bench.py --scale 22 --check-scale 14 --out $O/transpose_s22.json > $O/transpose_s22.log 2>&1; echo "transpose rc=$?"; tail -2 $O/transpose_s22.log | cut -c1-900
