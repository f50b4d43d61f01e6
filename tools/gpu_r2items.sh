# tools/gpu_r2items2.sh : round 2 -- item lists in owner order inside 256-owner stretches: step time and DRAM bytes of the walk kernels
set -x
O=gpurun_out/r2items2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_seam.py -m gpu -x -q -k "masked_dot or golden or tri_demo" 2>&1 | tail -3 > $O/pytest_gpu_dot.log
cat $O/pytest_gpu_dot.log
timeout 600 python tools/ab_tri.py --scale 22 --reps 4 --only default,nostreams --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
cut -c1-330 $O/ab_tri_s22.log | tail -3
timeout 300 python bench.py --slice-of 8 --slice-rank 0 --steps 5 --no-cpu --no-e2e --no-api --no-secondary > $O/bench_tri_slice0of8.json 2> $O/bench_tri_slice0of8.err
python tools/show_bench.py $O/bench_tri_slice0of8.json | cut -c1-200
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k regex:'dotr_kernel|dotr_warp_kernel|dot_kernel|dotg_items_all' -c 40 --csv --log-file $O/tri_s22_dram.csv python bench.py --workload tri --scale 22 --steps 1 --warmup 0 --no-cpu --no-e2e --no-api --no-secondary > $O/tri_s22_dram.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open("gpurun_out/r2items2/tri_s22_dram.csv")))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if len(r)>5 and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg.setdefault((d["ID"],d["Kernel Name"][:40]),{})[d["Metric Name"]]=(d["Metric Value"],d["Metric Unit"])
tot=0
for k,v in agg.items():
    print(k, {m:x for m,x in v.items()})
P
