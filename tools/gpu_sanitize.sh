# tools/gpu_sanitize.sh TAG : compute-sanitizer over the tests that drive the newest kernels
set -x
TAG=${1:-sanitize}
O=gpurun_out/$TAG; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_seam.py -x -q -k "hubs or edge_cases or typecast or golden or tri_demo" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_seam.py -x -q -k "masked_dot_hubs and True-PLUS-TIMES-INT64" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
tail -n 12 $O/memcheck.log $O/racecheck.log; cat $O/pytest.log
