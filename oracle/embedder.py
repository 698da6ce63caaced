"""CPU oracle of the FaceEmbedder flow — TEST INFRASTRUCTURE ONLY (see oracle/backbone.py header).

Restates reference face_embedder.py:112-182 over oracle.backbone: per-image preprocess,
torch fp32 eager forward in chunks of `batch_size`, model output -> numpy, optional
e / (||e|| + 1e-8).  This is also the "reference CPU path" bench.py times.
"""
import numpy as np
import torch

from . import backbone, preprocess as pp


class OracleEmbedder:
    def __init__(self, architecture="ir_101", model_type="adaface", state_dict=None, seed=0):
        self.architecture, self.model_type = architecture, model_type
        self.layout = "adaface" if model_type == "adaface" else "iresnet"
        self.sd = state_dict if state_dict is not None else backbone.random_state_dict(architecture, self.layout, seed)

    def extract_embeddings_batch(self, face_images, normalize=True, batch_size=32) -> np.ndarray:
        if len(face_images) == 0:
            return np.array([])
        outs = []
        for i in range(0, len(face_images), batch_size):
            x = torch.from_numpy(pp.preprocess_batch(face_images[i:i + batch_size], self.model_type))
            y = backbone.forward(self.sd, x, self.architecture, self.layout)
            outs.append((y[0] if self.layout == "adaface" else y).numpy())
        e = np.vstack(outs)
        if normalize:
            e = e / (np.linalg.norm(e, axis=1, keepdims=True) + 1e-8)
        return e

    def extract_embedding(self, face_image, normalize=True) -> np.ndarray:
        return self.extract_embeddings_batch([face_image], normalize)[0]
