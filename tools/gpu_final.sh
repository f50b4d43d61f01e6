# tools/gpu_final.sh TAG : the measurement set that profiles/<TAG>/ is made of
set -x
TAG=${1:-r1}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/gpu.csv
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest_gpu.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
python bench.py > $O/bench_tri_s22.json 2> $O/bench_tri_s22.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_tri_s22_reference.json 2> $O/bench_tri_s22_reference.err
python bench.py --workload tri --scale 20 > $O/bench_tri_s20.json 2> $O/bench_tri_s20.err
python bench.py --workload sssp > $O/bench_sssp_s22.json 2> $O/bench_sssp_s22.err
python bench.py --workload bfs --steps 3 > $O/bench_bfs_s22.json 2> $O/bench_bfs_s22.err
python bench.py --workload bfs --bfs-dir pull --steps 3 --no-cpu > $O/bench_bfspull_s22.json 2> $O/bench_bfspull_s22.err
python bench.py --workload spgemm --steps 3 > $O/bench_spgemm_er20.json 2> $O/bench_spgemm_er20.err
python bench.py --workload spgemm_rmat --scale 16 --steps 3 > $O/bench_spgemm_rmat16.json 2> $O/bench_spgemm_rmat16.err
python bench.py --workload spgemm_rmat --scale 18 --ef 8 --steps 2 --no-cpu > $O/bench_spgemm_rmat18.json 2> $O/bench_spgemm_rmat18.err
GB200_DOTG_ISO=0 python bench.py --steps 3 --no-cpu --no-e2e > $O/bench_tri_s22_valued.json 2> $O/bench_tri_s22_valued.err
tools/prof.sh $O tri_s22 dotg_kernel 4 --workload tri --scale 22
tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
tools/prof.sh $O sssp_s22 spmv_stream 1 --workload sssp
tools/launches.sh $O/sssp_s22_launches.csv --workload sssp
tools/prof.sh $O bfs_s22 saxpyv 12 --workload bfs
tools/launches.sh $O/bfs_s22_launches.csv --workload bfs
tools/prof.sh $O spgemm_rmat16 'saxpy_heavy|saxpy_light|sym_bitmap|heavy_mark|heavy_rank' 40 --workload spgemm_rmat --scale 16
tools/launches.sh $O/spgemm_rmat16_launches.csv --workload spgemm_rmat --scale 16
rm -f $O/*.ncu-rep $O/plain_*.log $O/ncu_*.log
du -sh $O; tail -n 3 $O/*.err; cat $O/pytest_gpu.log $O/smoke.log | tail -8
