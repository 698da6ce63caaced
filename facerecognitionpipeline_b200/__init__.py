"""facerecognitionpipeline_b200 — B200-native (sm_100a) embed-and-match hot path behind the
tuoasty/FaceRecognitionPipeline `face_embedder` / `gallery_manager` / `face_matcher` API."""

__version__ = "0.1.0"
