set -x
FRB_MULTI=1 timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -8
for m in 0 1; do
echo "MULTI=$m"; FRB_MULTI=$m timeout 300 python tools/bench_small.py 2>&1 | tail -5
done
