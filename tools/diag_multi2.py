#!/usr/bin/env python
"""Order dependence check: small IR-50 embeds first, then IR-101 at batch 1024 in the same process/context."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
from facerecognitionpipeline_b200 import weights
tag = sys.argv[1]
rng = np.random.default_rng(5)
crops = [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(1024)]
fe50 = FaceEmbedder(architecture="ir_50", model_type="adaface", state_dict=weights.random_init_state_dict("ir_50", "adaface", seed=1), max_batch=32)
for i in range(8): fe50.extract_embeddings_batch(crops[i:i + 1])
fe50.extract_embeddings_batch(crops[:32])
fe = FaceEmbedder(architecture="ir_101", model_type="adaface", state_dict=weights.random_init_state_dict("ir_101", "adaface", seed=0), max_batch=1024)
outs = [fe.extract_embeddings_batch(crops) for _ in range(4)]
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/diag2_{tag}.npy", np.stack(outs))
print(tag, "repeatable:", [bool(np.array_equal(outs[0], o)) for o in outs[1:]])
