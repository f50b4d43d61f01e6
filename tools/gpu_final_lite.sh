# tools/gpu_final_lite.sh TAG : re-measure the headline lines only (tests, smoke, tri, sssp + the tri ncu capture)
set -x
TAG=${1:-r1}
O=gpurun_out/$TAG; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
python bench.py > $O/bench_tri_s22.json 2> $O/bench_tri_s22.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_tri_s22_reference.json 2> $O/bench_tri_s22_reference.err
python bench.py --workload tri --scale 20 > $O/bench_tri_s20.json 2> $O/bench_tri_s20.err
python bench.py --workload sssp > $O/bench_sssp_s22.json 2> $O/bench_sssp_s22.err
GB200_DOTG_ISO=0 python bench.py --steps 3 --no-cpu --no-e2e > $O/bench_tri_s22_valued.json 2> $O/bench_tri_s22_valued.err
tools/prof.sh $O tri_s22 dotg_kernel 4 --workload tri --scale 22
tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
tools/prof.sh $O sssp_s22 spmv_stream 1 --workload sssp
rm -f $O/*.ncu-rep $O/plain_*.log $O/ncu_*.log
du -sh $O; tail -n 3 $O/*.err; cat $O/pytest_gpu.log $O/smoke.log | tail -6
