# tools/gpu_r2final3.sh : round 2 -- the final code (item lists ordered inside 256-owner stretches): masked-dot tests, A/B line, one rank of eight, DRAM bytes of the walk, the default bench line
set -x
O=gpurun_out/r2final3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 4 --only default --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -2
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
echo "rc=$?"; tail -3 $O/bench_default.err; grep real $O/bench_default.time; python tools/show_bench.py $O/bench_default.json 2>/dev/null | cut -c1-220
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k regex:'dotr_kernel|dotr_warp_kernel|dot_kernel|dotg_items_all' -c 40 --csv --log-file $O/tri_s22_dram.csv python bench.py --workload tri --scale 22 --steps 1 --warmup 0 --no-cpu --no-e2e --no-api --no-secondary > $O/tri_s22_dram.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2final3/tri_s22_dram.csv")))
hdr=None; rd=wr=0
for r in rows:
    if len(r)>5 and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if "items_all" in d["Kernel Name"]:
            if d["Metric Name"]=="gpu__time_duration.sum": print("items kernel ns", d["Metric Value"])
            continue
        if d["Metric Name"]=="dram__bytes_read.sum": rd+=float(d["Metric Value"])
        if d["Metric Name"]=="dram__bytes_write.sum": wr+=float(d["Metric Value"])
print("walk kernels DRAM read %.3f GB write %.3f GB total %.3f GB" % (rd/1e9, wr/1e9, (rd+wr)/1e9))
P
timeout 300 tools/launches.sh $O/tri_s22_launches.csv --workload tri --scale 22
