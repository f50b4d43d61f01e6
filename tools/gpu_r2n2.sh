# tools/gpu_r2n2.sh : round 2, two GPUs -- the in-library peer exchange (check + SSSP bench) and the tri line
set -x
O=gpurun_out/r2n2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_check.py --scale 18 > $O/peer_check.log 2>&1
echo "rc=$?"; tail -12 $O/peer_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload sssp --steps 10 --warmup 3 --no-cpu > $O/bench_sssp_n2.json 2> $O/bench_sssp_n2.err
echo "rc=$?"; tail -5 $O/bench_sssp_n2.err; cut -c1-900 $O/bench_sssp_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > $O/bench_tri_n2.json 2> $O/bench_tri_n2.err
echo "rc=$?"; tail -5 $O/bench_tri_n2.err; cut -c1-900 $O/bench_tri_n2.json
python tools/show_bench.py $O/bench_*.json 2>/dev/null | cut -c1-220
