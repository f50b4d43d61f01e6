# tools/gpu_r2f.sh : round 2 -- whole GPU suite; masked dot with vector records + faster bitmap build;
# fused shared-memory hash saxpy (GB200_SAXPY_HASH=0: the two-kernel path) on ER 2^20 and RMAT 16/18
set -x
O=gpurun_out/r2f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 400 python tools/ab_tri.py --scale 22 --reps 3 --only default,hub4096,hub2048,notiny,valued --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -8
for v in 1 0; do
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm --no-cpu --no-api --no-e2e --steps 5 > $O/bench_er20_hash$v.json 2> $O/bench_er20_hash$v.err
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm_rmat --scale 16 --no-cpu --no-api --no-e2e --steps 5 > $O/bench_rmat16_hash$v.json 2> $O/bench_rmat16_hash$v.err
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm_rmat --scale 18 --ef 8 --no-cpu --no-api --no-e2e --steps 3 > $O/bench_rmat18_hash$v.json 2> $O/bench_rmat18_hash$v.err
done
python tools/show_bench.py $O/bench_*_hash*.json 2>/dev/null | cut -c1-250
timeout 300 tools/launches.sh $O/er20_launches.csv --workload spgemm
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
du -sh $O
