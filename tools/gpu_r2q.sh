# tools/gpu_r2q.sh : round 2 -- hybrid product walk (long vectors of A on their own, short ones end to end), register sort in the warp-per-vector hash kernel
set -x
O=gpurun_out/r2q; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_seam.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
for w in "spgemm" "spgemm_rmat --scale 16" "spgemm_rmat --scale 18 --ef 8"; do
  n=$(echo $w | tr -d ' -'); 
  timeout 600 python bench.py --workload $w --steps 5 --no-e2e --no-api --no-secondary > $O/bench_$n.json 2> $O/bench_$n.err
  python tools/show_bench.py $O/bench_$n.json | cut -c1-200
done
timeout 300 tools/launches.sh $O/spgemm_er20_launches.csv --workload spgemm
timeout 300 tools/launches.sh $O/spgemm_rmat16_launches.csv --workload spgemm_rmat --scale 16
grep -c . $O/spgemm_er20_launches.csv
