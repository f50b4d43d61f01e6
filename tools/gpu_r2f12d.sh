# tools/gpu_r2f12d.sh : round 2 -- launch list of the final device transpose (scale 22; launches before the first 2368-block tr_vecof_kernel belong to the oracle-parity leg at scale 10)
O=gpurun_out/r2f12c; mkdir -p $O
timeout 42 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'tr_|scan_kernel|cast_kernel|nonempty_kernel|scatter_counts|hyper_pack|iso_kernel' -c 90 --csv --log-file $O/transpose_s22_final_launches.csv python tools/transpose_bench.py --scale 22 --check-scale 10 --reps 1 > $O/transpose_s22_final_ncu.log 2>&1
echo "ncu rc=$?"; wc -l $O/transpose_s22_final_launches.csv
