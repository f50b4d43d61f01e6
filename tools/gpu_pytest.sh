set -x
O=gpurun_out/r1_fin; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
cat $O/pytest.log
