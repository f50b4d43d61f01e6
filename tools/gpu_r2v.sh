# tools/gpu_r2v.sh : round 2 -- the k-truss loop as a caller of the masked saxpy, checked against host supports
set -x
O=gpurun_out/r2v; mkdir -p $O
timeout 600 python tools/ktruss.py --scale 18 --k 4 --check-scale 13 --out $O/ktruss_s18_k4.json > $O/ktruss_s18_k4.log 2>&1
echo "rc=$?"; tail -2 $O/ktruss_s18_k4.log | cut -c1-1200
