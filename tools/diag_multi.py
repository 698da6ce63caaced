#!/usr/bin/env python
"""Embeddings of the same seeded crops under the current FRB_* environment -> gpurun_out/diag_<tag>.npy"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
from facerecognitionpipeline_b200 import weights
tag, arch, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
sd = weights.random_init_state_dict(arch, "adaface", seed=0)
fe = FaceEmbedder(architecture=arch, model_type="adaface", state_dict=sd, max_batch=B)
rng = np.random.default_rng(5)
crops = [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(B)]
outs = [fe.extract_embeddings_batch(crops) for _ in range(3)]
np.save(f"gpurun_out/diag_{tag}.npy", np.stack(outs))
print(tag, "repeatable:", [bool(np.array_equal(outs[0], o)) for o in outs[1:]])
