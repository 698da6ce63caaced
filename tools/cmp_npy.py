#!/usr/bin/env python
"""cmp_npy.py a.npy b.npy ... : are the arrays bit-identical to the first one? (schedule experiments)"""
import sys

import numpy as np

ref = np.load(sys.argv[1])
for p in sys.argv[2:]:
    o = np.load(p)
    same = o.shape == ref.shape and np.array_equal(o, ref)
    print(p, "IDENTICAL" if same else f"DIFFERENT ({int((o != ref).any(-1).sum()) if o.shape == ref.shape else 'shape'} rows)")
