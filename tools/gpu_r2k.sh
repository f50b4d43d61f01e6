# tools/gpu_r2k.sh : round 2 -- what one rank of an 8-GPU triangle count does (slice 0 of 8 on one GPU): launch list
set -x
O=gpurun_out/r2k; mkdir -p $O
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
timeout 300 tools/launches.sh $O/tri_slice0of8_launches.csv --workload tri --scale 22 --slice-of 8 --slice-rank 0 --calibrate 0
GB200_TRACE=1 timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 2 --no-cpu --no-e2e --no-api --no-secondary --calibrate 0 2>&1 | tail -3 | cut -c1-300
