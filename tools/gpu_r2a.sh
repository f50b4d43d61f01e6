# tools/gpu_r2a.sh : round 2, first GPU session -- the flat masked-dot kernels against the round-1 ones
set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $O/gpu.csv
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 300 python tools/ab_tri.py --scale 22 --reps 3 --only default,old,valued,valued_old,bm_half,chunk512,chunk2048,hub4096,hub16384 --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -12
timeout 300 python bench.py --no-cpu > $O/bench_tri_s22.json 2> $O/bench_tri_s22.err
cut -c1-600 $O/bench_tri_s22.json
timeout 600 tools/prof.sh $O tri_s22 'dotf_kernel|dotg_kernel|dot_kernel' 8 --workload tri --scale 22
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
rm -f $O/plain_*.log $O/ncu_*.log
du -sh $O
