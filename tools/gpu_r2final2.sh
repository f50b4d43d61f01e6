# tools/gpu_r2final2.sh : round 2 -- the final code (trim searches on by default): default bench line, ncu capture + launch list of the masked-dot step, the masked-dot tests
set -x
O=gpurun_out/r2final2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
echo "rc=$?"; tail -3 $O/bench_default.err; grep real $O/bench_default.time; python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-220
timeout 600 tools/prof.sh $O tri_s22 'dotr_kernel|dotr_warp_kernel|dot_kernel' 7 --workload tri --scale 22
python tools/ncu_summary.py $O/tri_s22_raw.csv > $O/tri_s22_ncu_summary.txt 2>&1
grep -E "dram__bytes_read.sum |dram__bytes_write.sum " $O/tri_s22_ncu_summary.txt | cut -c1-150
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
rm -f $O/plain_*.log $O/ncu_*.log $O/*_source.csv.tmp
