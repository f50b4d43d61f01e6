# tools/gpu_ab.sh TAG : A/B of the masked-dot variants at scale 22 (bit-exact against the first
# variant, times per variant), then the masked-dot parity tests against the oracle / the reference
TAG=${1:-ab}
O=gpurun_out/$TAG; mkdir -p $O
timeout 60 python tools/ab_tri.py --scale 22 --reps 2 --out $O/ab_tri_s22.json > $O/ab_tri_s22.log 2>&1
tail -n 20 $O/ab_tri_s22.log | cut -c1-400
timeout 40 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py -m gpu -x -q \
    -k "masked_dot or tri_demo or golden or tricount or test_masked or hypersparse or ktruss" 2>&1 | tail -5 > $O/pytest_dot.log
cat $O/pytest_dot.log
