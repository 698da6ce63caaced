#!/bin/bash
# One GPU-box pass: the -m gpu suite, the default bench line, schedule A/B runs (bit-identity + timing).
#   gpurun --timeout 1700 -- 'bash tools/run_gpu_checks.sh <tag> [all|ab|abonly|summary]'   (summary: only digest gpurun_out/<tag>_bench*.log)
# Writes gpurun_out/<tag>_*.log; prints a short summary.
tag=${1:-run}; mode=${2:-all}
mkdir -p gpurun_out
if [ "$mode" != "abonly" ] && [ "$mode" != "summary" ]; then
  (timeout 1300 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/${tag}_pytest.log 2>&1
  tail -4 gpurun_out/${tag}_pytest.log
fi
if [ "$mode" != "summary" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2>&1
fi
if [ "$mode" = "ab" ] || [ "$mode" = "abonly" ]; then
  for m in 0 1; do for b in 8 256; do
    FRB_SLAB_MULTI=$m timeout 300 python tools/diag_multi.py ${tag}_sm${m}_$b ir_101 $b 2>&1 | tail -1
  done; done
  python tools/cmp_npy.py gpurun_out/diag_${tag}_sm0_8.npy gpurun_out/diag_${tag}_sm1_8.npy
  python tools/cmp_npy.py gpurun_out/diag_${tag}_sm0_256.npy gpurun_out/diag_${tag}_sm1_256.npy
  FRB_SLAB_MULTI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/${tag}_bench_sm1.log 2>&1
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.log")):
    try:
        d = json.loads([x for x in open(f) if x.startswith("{")][-1])
        pk = {k["kernel"][:28]: round(k["us_per_step"], 1) for k in d["roofline"]["per_kernel"]}
        print(f, "faces/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "embed_ms", round(d["embed_ms"], 3), "match_ms", round(d["match_ms"], 3),
              "c4", round(d.get("c4", {}).get("faces_per_s", 0)), "c3_ms", round(d["match_4096"]["ms"], 3))
        print("   ", pk)
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-600:])
PY
