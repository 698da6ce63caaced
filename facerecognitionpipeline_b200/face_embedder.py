"""Drop-in for the reference `face_embedder` module (face_embedder.py:26-225).

`FaceEmbedder` keeps the constructor, attributes (.device, .model_type, .architecture, .input_size)
and methods of the reference; the forward pass runs in libfrb200 (sm_100a tcgen05 kernels) through
the C ABI instead of torch-eager / onnxruntime.  There is no CPU fallback: constructing an embedder
without a B200 raises.

Differences that are visible and intended:
  * `state_dict=` (new, optional) lets callers pass weights directly (random-init for tests/bench;
    no checkpoint files are shipped with the reference either, .gitignore:15,18,33).
  * `batch_size` in `extract_embeddings_batch` is accepted but the device path chunks by
    `self.max_batch` (eval-mode results do not depend on the chunking).
  * ArcFace: the reference runs an ONNX export of insightface iresnet through onnxruntime
    (face_embedder.py:64-88).  Here the `.onnx` file's initializers are read directly (`onnx_import`, no onnx /
    onnxruntime package) and mapped onto the same device program an iresnet `.pth` state dict builds.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List

import numpy as np

from . import _native, templates, weights

SCRIPT_DIR = Path(__file__).resolve().parent
_PRETRAINED = Path(os.environ.get("FRB_PRETRAINED_DIR", SCRIPT_DIR / "pretrained"))

ADAFACE_MODELS = {
    "ir_50": str(_PRETRAINED / "adaface_ir50_ms1mv2.ckpt"),
    "ir_101": str(_PRETRAINED / "adaface_ir101_ms1mv3.ckpt"),
}
ARCFACE_MODELS = {
    "ir_50": str(_PRETRAINED / "arcface_ir50_ms1mv3.onnx"),
    "ir_101": str(_PRETRAINED / "arcface_ir101_ms1mv3.onnx"),
}


class _Device:
    """Minimal stand-in for the torch.device the reference exposes as `.device`
    (read by face_recognition_server.py:237)."""

    def __init__(self, index):
        self.type, self.index = "cuda", index

    def __str__(self):
        return f"cuda:{self.index}"

    __repr__ = __str__


class FaceEmbedder:
    def __init__(self, architecture="ir_101", model_path=None, model_type="adaface", device=None, state_dict=None,
                 max_batch: int = 256):
        index = 0
        if device is not None:
            index = getattr(device, "index", None)
            if index is None:
                index = int(str(device).split(":")[1]) if ":" in str(device) else 0
        self.device = _Device(index)
        self.model_type = model_type
        self.architecture = architecture
        self.max_batch = int(max_batch)

        if model_type == "adaface":
            layout, table = "adaface", ADAFACE_MODELS
            self.mean, self.std = 0.5, 0.5
        elif model_type == "arcface":
            layout, table = "iresnet", ARCFACE_MODELS
            self.mean, self.std = 127.5, 127.5
        else:
            raise ValueError(f"Unknown model_type: {model_type}. Must be 'adaface' or 'arcface'")
        if architecture not in table:
            raise ValueError(f"Unknown architecture: {architecture}. Available: {list(table.keys())}")

        if state_dict is None:
            if model_path is None:
                model_path = table[architecture]
            if not os.path.exists(model_path):
                raise FileNotFoundError(f"Model file not found at: {model_path}")
            print(f"Loading {model_type} model ({architecture}) from {model_path}...")
            if str(model_path).lower().endswith(".onnx"):
                from . import onnx_import
                if layout != "iresnet":
                    raise ValueError("an .onnx model file holds the ArcFace (insightface iresnet) export: use model_type='arcface'")
                state_dict, found = onnx_import.onnx_to_iresnet_state_dict(model_path)
                if found != architecture:
                    raise ValueError(f"{model_path} holds an {found} backbone, architecture={architecture!r} was requested")
            else:
                state_dict = weights.load_checkpoint_state_dict(model_path)
        self._program = weights.build_program(state_dict, architecture, layout)
        self._ctx = _native.default_context(index)
        with self._ctx.lock:
            self._program.load_into(self._ctx)
            self._loaded_gen = self._ctx.backbone_generation()
        self._flags = _native.FRB_EMBED_L2 if layout == "adaface" else 0
        self.input_size = (112, 112)
        self.is_onnx = str(model_path or "").lower().endswith(".onnx")   # informational, as the reference's attribute
        self.model = self  # the reference exposes `.model`; here the embedder is the model handle
        print(f"{model_type} model loaded on {self.device} (libfrb200, sm_100a)")

    # ------------------------------------------------------------------ preprocessing
    def preprocess(self, face_image: np.ndarray) -> np.ndarray:
        """Host restatement for API completeness: returns the [1,3,112,112] float32 BGR tensor the
        reference feeds its model (face_embedder.py:93-110).  The device path does NOT use this; it
        consumes uint8 crops directly (frb_preprocess_u8)."""
        import cv2
        if face_image.shape[:2] != self.input_size:
            face_image = cv2.resize(face_image, self.input_size, interpolation=cv2.INTER_LINEAR)
        bgr = face_image[:, :, ::-1]
        if self.model_type == "adaface":
            x = (bgr / 255.0 - self.mean) / self.std
        else:
            x = (bgr - self.mean) / self.std
        return np.expand_dims(x.transpose(2, 0, 1), 0).astype(np.float32)

    def _to_device_sizes(self, face_images: List[np.ndarray]):
        """Group crops into uint8 stacks the device preprocess accepts (112 direct, 224 -> 2x2 box mean,
        which is what cv2.resize INTER_LINEAR computes at exactly 2x).  Any other size goes through
        cv2.resize on the host exactly as the reference does (face_embedder.py:94-96)."""
        import cv2
        groups = {112: [], 224: []}
        order = []
        for img in face_images:
            img = np.asarray(img)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError(f"expected HxWx3 RGB uint8 image, got shape {img.shape}")
            if img.dtype != np.uint8:
                img = img.astype(np.uint8)
            hw = img.shape[:2]
            if hw == (224, 224):
                s = 224
            else:
                s = 112
                if hw != (112, 112):
                    img = cv2.resize(img, self.input_size, interpolation=cv2.INTER_LINEAR)
            order.append((s, len(groups[s])))
            groups[s].append(np.ascontiguousarray(img))
        return groups, order

    def _ensure_loaded(self):
        """Reload this embedder's program when anything else was loaded into the shared context since
        (`frb_backbone_generation` counts the loads).  Call with the context's lock held."""
        if self._loaded_gen != self._ctx.backbone_generation():
            self._program.load_into(self._ctx)
            self._loaded_gen = self._ctx.backbone_generation()

    def _embed_u8(self, stack: np.ndarray, S: int, normalize: bool) -> np.ndarray:
        n = len(stack)
        out = np.empty((n, 512), np.float32)
        flags = self._flags | (_native.FRB_EMBED_RENORM if normalize else 0)
        chunks = [np.ascontiguousarray(stack[i:i + self.max_batch]) for i in range(0, n, self.max_batch)]
        with self._ctx.lock:   # check + calls are one unit: no other embedder may swap the weights in between
            self._ensure_loaded()
            for j, chunk in enumerate(chunks):
                if j + 1 < len(chunks):   # the next chunk's crops travel while this one computes
                    self._ctx.frb_prefetch_host(chunks[j + 1].ctypes.data, len(chunks[j + 1]), S)
                dst = out[j * self.max_batch:j * self.max_batch + len(chunk)]
                self._ctx.frb_embed_host(chunk.ctypes.data, len(chunk), S, flags, dst.ctypes.data, None)
        return out

    # ------------------------------------------------------------------ embedding API
    def extract_embedding(self, face_image: np.ndarray, normalize=True) -> np.ndarray:
        return self.extract_embeddings_batch([face_image], normalize=normalize)[0]

    def extract_embeddings_batch(self, face_images: List[np.ndarray], normalize=True, batch_size=32) -> np.ndarray:
        if len(face_images) == 0:
            return np.array([])
        groups, order = self._to_device_sizes(face_images)
        embs = {}
        for s, imgs in groups.items():
            if imgs:
                embs[s] = self._embed_u8(np.stack(imgs), s, normalize)
        return np.stack([embs[s][j] for s, j in order])

    # ------------------------------------------------------------------ small helpers (host, as reference)
    def compute_similarity(self, embedding1: np.ndarray, embedding2: np.ndarray) -> float:
        a = embedding1 / (np.linalg.norm(embedding1) + 1e-8)
        b = embedding2 / (np.linalg.norm(embedding2) + 1e-8)
        return np.dot(a, b)

    def compute_similarity_batch(self, embedding: np.ndarray, gallery_embeddings: np.ndarray) -> np.ndarray:
        q = embedding / (np.linalg.norm(embedding) + 1e-8)
        g = gallery_embeddings / (np.linalg.norm(gallery_embeddings, axis=1, keepdims=True) + 1e-8)
        return np.dot(g, q)

    def aggregate_embeddings(self, embeddings: np.ndarray, method="mean") -> np.ndarray:
        return templates.embedder_template(embeddings, method)
