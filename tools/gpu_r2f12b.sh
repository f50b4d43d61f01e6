# tools/gpu_r2f12b.sh : round 2, final code -- the default bench line (with the neighbours entry)
set -x
O=gpurun_out/r2f12; mkdir -p $O
( time timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
echo "rc=$?"; tail -3 $O/bench_default.err; grep real $O/bench_default.time; python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-260
python - <<'P'
import json
l=json.loads(open("gpurun_out/r2f12/bench_default.json").read().strip().splitlines()[-1])
print("parity", l.get("parity")); print("neighbours", json.dumps(l.get("neighbours"))[:700])
P
# the launch list of the device transpose (scale 22; the first launches belong to the oracle-parity leg at scale 10)
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'tr_|scan_kernel|cast_kernel|nonempty_kernel|scatter_counts|hyper_pack' -c 120 --csv --log-file $O/transpose_s22_launches.csv python tools/transpose_bench.py --scale 22 --check-scale 10 --reps 1 > $O/transpose_s22_ncu.log 2>&1
echo "ncu rc=$?"; tail -30 $O/transpose_s22_launches.csv | cut -c1-200 | awk -F'","' '{print $5" | "$(NF-2)" | "$NF}' | tail -24
