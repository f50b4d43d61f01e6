# tools/gpu_r2b.sh : round 2 -- the restructured bench.py (parity gate, t_api, secondary, reference arm)
set -x
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python bench.py --scale 16 --secondary-scale 16 --steps 2 > $O/bench_small.json 2> $O/bench_small.err
echo "rc=$?"; tail -5 $O/bench_small.err; cut -c1-1500 $O/bench_small.json
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "rc=$?"; tail -5 $O/bench_default.err; cut -c1-3000 $O/bench_default.json
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
echo "rc=$?"; tail -3 $O/bench_reference.err; cut -c1-1200 $O/bench_reference.json
nproc; free -g | head -2
