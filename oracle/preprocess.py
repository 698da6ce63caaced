"""CPU oracle of FaceEmbedder.preprocess — TEST INFRASTRUCTURE ONLY (see oracle/backbone.py header).

Follows reference face_embedder.py:93-110 step by step with the same library calls:
resize to 112x112 with cv2.INTER_LINEAR when the crop is another size (:94-96), RGB->BGR (:99/:106),
adaface `(x / 255.0 - 0.5) / 0.5` in float64 then float32 (:100-101), arcface `(x - 127.5) / 127.5`
then float32 (:107-110), HWC->CHW, leading batch axis.
"""
import cv2
import numpy as np


def preprocess(face_image: np.ndarray, model_type: str = "adaface") -> np.ndarray:
    """RGB uint8 HxWx3 -> float32 [1,3,112,112] (BGR planes)."""
    if face_image.shape[:2] != (112, 112):
        face_image = cv2.resize(face_image, (112, 112), interpolation=cv2.INTER_LINEAR)
    bgr = face_image[:, :, ::-1]
    if model_type == "adaface":
        x = (bgr / 255.0 - 0.5) / 0.5
    elif model_type == "arcface":
        x = (bgr - 127.5) / 127.5
    else:
        raise ValueError(f"Unknown model_type: {model_type}. Must be 'adaface' or 'arcface'")
    return np.expand_dims(x.transpose(2, 0, 1), axis=0).astype(np.float32)


def preprocess_batch(face_images, model_type: str = "adaface") -> np.ndarray:
    return np.concatenate([preprocess(im, model_type) for im in face_images], axis=0)
