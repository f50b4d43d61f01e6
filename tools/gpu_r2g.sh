# tools/gpu_r2g.sh : round 2 -- whole GPU suite; fused hash saxpy with bitmap-ranked output, A/B
set -x
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
for v in 1 0; do
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm --no-cpu --no-api --no-e2e --steps 5 > $O/bench_er20_hash$v.json 2> $O/bench_er20_hash$v.err
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm_rmat --scale 16 --no-cpu --no-api --no-e2e --steps 5 > $O/bench_rmat16_hash$v.json 2> $O/bench_rmat16_hash$v.err
  GB200_SAXPY_HASH=$v timeout 300 python bench.py --workload spgemm_rmat --scale 18 --ef 8 --no-cpu --no-api --no-e2e --steps 3 > $O/bench_rmat18_hash$v.json 2> $O/bench_rmat18_hash$v.err
done
python tools/show_bench.py $O/bench_*_hash*.json 2>/dev/null | cut -c1-250
GB200_SAXPY_HASH=1 timeout 300 tools/launches.sh $O/rmat16_hash1_launches.csv --workload spgemm_rmat --scale 16
GB200_SAXPY_HASH=0 timeout 300 tools/launches.sh $O/rmat16_hash0_launches.csv --workload spgemm_rmat --scale 16
du -sh $O
