mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_embed.py tests/test_gpu_e2e.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
for r in 1 2; do for L in old new; do
  if [ $L = old ]; then export FRB_LIBRARY=$PWD/facerecognitionpipeline_b200/libfrb200_old.so; else unset FRB_LIBRARY; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-c4 > gpurun_out/r3j_$L$r.log 2>&1
  python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/r3j_$L$r.log") if x.startswith("{")][-1])
print("$L$r", round(d["value"]), "embed", round(d["embed_ms"],3), "match", round(d["match_ms"],3), "c3", round(d["match_4096"]["ms"],3))
PY
done; done
unset FRB_LIBRARY
ncu --metrics gpu__time_duration.sum --clock-control none -c 180 --csv --log-file gpurun_out/r3j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r3j_ll.log 2>&1; echo ll rc=$?
