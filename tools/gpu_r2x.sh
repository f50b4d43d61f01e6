# tools/gpu_r2x.sh : round 2 -- where the non-kernel time of the saxpy multiplies goes (GB200_TRACE: trips of the block cache to the driver)
set -x
O=gpurun_out/r2x; mkdir -p $O
GB200_TRACE=1 timeout 600 python tools/ktruss.py --scale 18 --k 4 --check-scale 0 > $O/ktruss_trace.log 2>&1
grep "gb200_AxB_device" $O/ktruss_trace.log | cut -c1-200
GB200_TRACE=1 timeout 600 python bench.py --workload spgemm_rmat --scale 16 --steps 6 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_rmat16.json 2> $O/bench_rmat16.err
grep "gb200_AxB_device" $O/bench_rmat16.err | cut -c1-200 | tail -10
timeout 600 python tools/ktruss.py --scale 18 --k 4 --check-scale 13 --out $O/ktruss_s18_k4.json > $O/ktruss_s18_k4.log 2>&1
tail -1 $O/ktruss_s18_k4.log | cut -c1-700
