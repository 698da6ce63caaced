for cfg in "1 1" "2 1" "1 0"; do set -- $cfg
if [ "$2" = "1" ]; then export FRB_DF_NOFENCE=1; else unset FRB_DF_NOFENCE; fi
FRB_DATAFLOW=$1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r1m_bench.log 2>&1
tail -1 gpurun_out/r1m_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH df=$1 nofence=$2', d['value'], d['embed_ms'], d['match_ms'], d['clocks'])" || tail -5 gpurun_out/r1m_bench.log
done
