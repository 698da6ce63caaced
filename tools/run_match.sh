for sl in 5 9 14 19 37; do echo "slices=$sl (P=4096)"; FRB_MATCH_SLICES=$sl python tools/bench_match.py 4096; done
for sl in 19 37 56 74; do echo "slices=$sl (P=1024)"; FRB_MATCH_SLICES=$sl python tools/bench_match.py 1024; done
for sl in 74 148 222; do echo "slices=$sl (P=256)"; FRB_MATCH_SLICES=$sl python tools/bench_match.py 256; done
