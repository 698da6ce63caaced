"""FaceQualityFilter / FaceProcessor.process_numpy ordering / server best-frame selection against outputs of the
reference's own classes (tests/golden/make_golden_r2.py -> r2_cases.npz).  The filter and the host best-frame rule are
pure host code (CPU tests); process_numpy and the batched device arg-max need the GPU."""
import ast
import ctypes as C
import os

import numpy as np
import pytest

from facerecognitionpipeline_b200.face_recognition import FaceQualityFilter

KEYS = ["det_score", "face_size", "yaw", "pitch", "roll", "blur_score"]


@pytest.fixture(scope="module")
def r2(golden_dir):
    return np.load(os.path.join(golden_dir, "r2_cases.npz"))


def _faces(z, prefix):
    return [dict(bbox=z[prefix + "/bbox"][i], landmarks=z[prefix + "/landmarks"][i], det_score=z[prefix + "/det_score"][i])
            for i in range(len(z[prefix + "/bbox"]))]


def test_quality_filter_matches_reference_outputs(r2):
    faces = _faces(r2, "quality")
    crops = r2["quality/crops"][r2["quality/crop_of"]]
    for ci, cfg_repr in enumerate(r2["quality/configs"]):
        qf = FaceQualityFilter(**ast.literal_eval(str(cfg_repr)))
        for i, (f, c) in enumerate(zip(faces, crops)):
            ok, m = qf.is_valid(f, c)
            assert ok == bool(r2[f"quality/{ci}/valid"][i])
            assert [k in m for k in KEYS] == r2[f"quality/{ci}/present"][i].tolist()      # same early-exit point
            for j, k in enumerate(KEYS):
                if k in m:
                    assert float(m[k]) == r2[f"quality/{ci}/values"][i, j], (ci, i, k)     # bit-identical metrics


def test_best_frame_host_rule(r2):
    from facerecognitionpipeline_b200.face_matcher import best_frame_index
    seg, det, blur = r2["bestframe/seg"], r2["bestframe/det"], r2["bestframe/blur"]
    for t in range(len(seg) - 1):
        sl = slice(seg[t], seg[t + 1])
        idx, ready = best_frame_index(det[sl], blur[sl])
        assert idx == r2["bestframe/best"][t] and ready == bool(r2["bestframe/gate"][t])


@pytest.mark.gpu
def test_best_frames_on_device(ctx, r2):
    import torch
    dev = torch.device("cuda", 0)
    seg, det, blur = (torch.from_numpy(r2["bestframe/" + k]).to(dev) for k in ("seg", "det", "blur"))
    T = len(seg) - 1
    out_i = torch.empty((T,), dtype=torch.int64, device=dev)
    out_q = torch.empty((T,), dtype=torch.float64, device=dev)
    out_r = torch.empty((T,), dtype=torch.uint8, device=dev)
    ctx.frb_best_frames(det.data_ptr(), blur.data_ptr(), seg.data_ptr(), T, 0.6, out_i.data_ptr(), out_q.data_ptr(),
                        out_r.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out_i.cpu().numpy(), r2["bestframe/best"])
    assert np.array_equal(out_r.cpu().numpy().astype(bool), r2["bestframe/gate"])
    d, b = r2["bestframe/det"], r2["bestframe/blur"]
    q = d * np.minimum(b / 100.0, 1.0)
    s = r2["bestframe/seg"]
    assert np.array_equal(out_q.cpu().numpy(), np.array([q[s[t] + r2["bestframe/best"][t]] for t in range(T)]))
    from facerecognitionpipeline_b200.face_matcher import best_frames_batch
    bi, br = best_frames_batch(r2["bestframe/det"], r2["bestframe/blur"], r2["bestframe/seg"])
    assert np.array_equal(bi, r2["bestframe/best"]) and np.array_equal(br, r2["bestframe/gate"])


@pytest.mark.gpu
@pytest.mark.parametrize("S", [112, 224])
def test_process_numpy_order_and_flags(r2, S):
    from facerecognitionpipeline_b200.face_recognition import FaceProcessor
    dets = _faces(r2, "process")

    class Det:
        def detect(self, img):
            return [dict(f) for f in dets]

    for ci in (0, 1):
        cfg = ast.literal_eval(str(r2[f"process/config{ci}"]))
        fp = FaceProcessor(output_size=S, quality_filter_config=cfg, detector=Det())
        for ra, tag in ((True, "all"), (False, "best")):
            res = fp.process_numpy(r2["process/frame"], return_all=ra)
            order = [next(i for i, f in enumerate(dets) if np.array_equal(f["bbox"], r["bbox"])) for r in res]
            base = f"process/S{S}/c{ci}/{tag}"
            assert order == r2[base + "/order"].tolist()
            assert [r["is_valid"] for r in res] == r2[base + "/valid"].astype(bool).tolist()
            blur = [r["quality_metrics"].get("blur_score", -1.0) for r in res]
            assert blur == r2[base + "/blur"].tolist()                 # bit-identical: the device warp equals cv2's
            if ra and S == 112 and ci == 0:
                assert np.array_equal(np.stack([r["aligned_face"] for r in res]), r2["process/aligned112"])
