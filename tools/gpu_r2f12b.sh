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
