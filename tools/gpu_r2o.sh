# tools/gpu_r2o.sh : round 2 -- side streams + 16-byte vector records; ncu of the set-up kernels
set -x
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py -m gpu -x -q -k "masked_dot or golden or tri_demo" 2>&1 | tail -4 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 3 --only default,nostreams,trim1,nostreams_trim1,valued --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-420 $O/ab_tri_s22.log | tail -6
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
timeout 600 tools/prof.sh $O setup 'dotg_classify|dotg_scatter|dot_gather|expand_vec|scan_kernel' 8 --workload tri --scale 22
python tools/ncu_summary.py $O/setup_raw.csv > $O/setup_summary.txt 2>&1; head -c 3000 $O/setup_summary.txt
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
