# tools/gpu_demo.sh : the reference's own demo binaries under LD_PRELOAD of the shim
set -x
O=gpurun_out/r1_demo; mkdir -p $O
python -m pytest tests/test_demo_programs.py -m gpu -x -q 2>&1 | tail -25 > $O/pytest_demo.log
cat $O/pytest_demo.log
