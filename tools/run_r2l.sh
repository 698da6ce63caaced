for pf in 0 1 2 4 8; do echo "PREFETCH=$pf"; FRB_MATCH_PREFETCH=$pf timeout 200 python tools/bench_match.py 256 512 2>&1 | tail -2; done
