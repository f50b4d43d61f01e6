# tools/gpu_r2u.sh : round 2 -- launch list of the valued masked dot (pattern-only operands sent through the valued kernels)
set -x
O=gpurun_out/r2u; mkdir -p $O
GB200_DOTG_ISO=0 GB200_DOT_STREAMS=0 timeout 300 tools/launches.sh $O/tri_s22_valued_launches.csv --workload tri --scale 22
grep -c . $O/tri_s22_valued_launches.csv
