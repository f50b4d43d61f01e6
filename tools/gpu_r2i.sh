# tools/gpu_r2i.sh : round 2 -- capture of the current masked-dot kernels (ncu --set full), launch lists of
# the four bench workloads, the default bench line and the whole GPU suite
set -x
O=gpurun_out/r2i; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/gpu.csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "rc=$?"; tail -3 $O/bench_default.err; python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-220
timeout 600 tools/prof.sh $O tri_s22 'dotr_kernel|dotr_warp_kernel|dot_kernel' 7 --workload tri --scale 22
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
timeout 400 tools/prof.sh $O sssp_s22 'spmv_stream_kernel' 1 --workload sssp --scale 22
timeout 300 tools/launches.sh $O/sssp_s22_launches.csv --workload sssp --scale 22
timeout 300 tools/launches.sh $O/bfs_s22_launches.csv --workload bfs --scale 22
timeout 300 tools/launches.sh $O/spgemm_rmat16_launches.csv --workload spgemm_rmat --scale 16
timeout 300 tools/launches.sh $O/spgemm_er20_launches.csv --workload spgemm
rm -f $O/plain_*.log $O/ncu_*.log $O/*_source.csv.tmp
du -sh $O
