set -x
for dbg in 0 1 3 7 15; do
echo "TS_DEBUG=$dbg"
FRB_TAIL_SPLIT=1 FRB_TS_DEBUG=$dbg timeout 100 python tools/microbench_gemm.py one 256 14 256 256 2>&1 | tail -1
FRB_TAIL_SPLIT=1 FRB_TS_DEBUG=$dbg timeout 100 python tools/microbench_gemm.py one 97 14 256 256 2>&1 | tail -1
done
for f in 2 4; do
FRB_TAIL_SPLIT=1 FRB_TS_FORCE=$f timeout 100 python tools/microbench_gemm.py one 256 14 256 256 2>&1 | tail -1
done
FRB_TAIL_SPLIT=0 timeout 100 python tools/microbench_gemm.py one 256 14 256 256 2>&1 | tail -1
