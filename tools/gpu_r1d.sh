set -x
mkdir -p gpurun_out/r1d
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1d/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d/smoke.log 2>&1
python bench.py > gpurun_out/r1d/bench_default.json 2> gpurun_out/r1d/bench_default.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1d/bench_ref.json 2> gpurun_out/r1d/bench_ref.err
nvidia-smi > gpurun_out/r1d/smi.log; nproc >> gpurun_out/r1d/smi.log; free -g >> gpurun_out/r1d/smi.log
tail -3 gpurun_out/r1d/*.err gpurun_out/r1d/pytest.log gpurun_out/r1d/smoke.log
