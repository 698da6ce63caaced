#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE itself (run in the build container,
where /root/reference is mounted; the GPU box never sees the reference):

  gallery_backups.npz   : the shipped full-embedding backups (gallery/backups/*.json + backups/*.json):
                          per file 23 students x (8 x 512 embeddings, 512 template)  -> pins filter+mean+renorm
  aggregate_cases.npz   : outputs of the reference's own GalleryManager._aggregate_embeddings /
                          _filter_quality_embeddings / _remove_outliers on seeded inputs (all 3 methods)
  search_cases.npz      : outputs of the reference's own GalleryManager.search on the adaface_ir_101
                          backup gallery (23 identities) for seeded probes
  track_001.json        : the recorded per-frame matches + final decision of
                          output/camera_captures/track_001/recognition_result.json -> pins _aggregate_matches
  preprocess_cases.npz  : seeded crops -> the float32 tensors FaceEmbedder.preprocess's arithmetic yields
                          (face_embedder.py:93-110 executed line by line; the module itself cannot be
                          imported: `import net` fails)
"""
import contextlib
import glob
import io
import json
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        import gallery_manager as ref_gm  # the reference's own module

    # ---- shipped backups
    files = sorted(glob.glob(f"{REF}/gallery/backups/*.json")) + sorted(glob.glob(f"{REF}/backups/*.json"))
    pack = {}
    for fp in files:
        d = json.load(open(fp))
        tag = os.path.basename(fp).replace(".json", "")
        sids = list(d["students"].keys())
        pack[tag + "/ids"] = np.array(sids)
        pack[tag + "/names"] = np.array([d["students"][s]["name"] for s in sids])
        pack[tag + "/emb"] = np.array([d["students"][s]["embeddings"] for s in sids], np.float32)
        pack[tag + "/tpl"] = np.array([d["students"][s]["template_embedding"] for s in sids], np.float32)
    np.savez_compressed(os.path.join(OUT, "gallery_backups.npz"), **pack)

    # ---- aggregation through the reference's own class
    rng = np.random.default_rng(1234)
    agg = {}
    with contextlib.redirect_stdout(io.StringIO()):
        case = 0
        for method in ("mean", "median", "weighted_mean", "bogus"):
            gm = ref_gm.GalleryManager(gallery_path="/tmp/frb_golden/none.pkl", aggregation_method=method)
            for n, spread in [(1, 0.1), (2, 0.3), (5, 0.15), (8, 0.25), (8, 0.9), (16, 0.45), (40, 0.35)]:
                base = rng.standard_normal(512)
                e = base[None, :] + spread * np.sqrt(512) * 0.1 * rng.standard_normal((n, 512)) * (1 + 3 * (rng.random((n, 1)) < 0.25))
                e = (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)
                agg[f"c{case}/method"] = np.array(method)
                agg[f"c{case}/in"] = e
                agg[f"c{case}/out"] = np.asarray(gm._aggregate_embeddings(e.copy()))
                agg[f"c{case}/filtered"] = np.asarray(gm._filter_quality_embeddings(e.copy()))
                agg[f"c{case}/outliers"] = np.asarray(gm._remove_outliers(e.copy()))
                case += 1
        agg["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(OUT, "aggregate_cases.npz"), **agg)

    # ---- search through the reference's own class on the shipped 23-identity gallery
    tag = "adaface_ir_101_backup_20251202_084813"
    with contextlib.redirect_stdout(io.StringIO()):
        gm = ref_gm.GalleryManager(gallery_path="/tmp/frb_golden/none2.pkl")
        for sid, name, emb, tpl in zip(pack[tag + "/ids"], pack[tag + "/names"], pack[tag + "/emb"], pack[tag + "/tpl"]):
            now = "2025-01-01T00:00:00"
            gm.students[str(sid)] = ref_gm.StudentRecord(str(sid), str(name), emb, tpl, len(emb), now, now, {})
        probes, exp_ids, exp_scores = [], [], []
        for i in range(64):
            s = i % 23
            q = pack[tag + "/emb"][s][i % 8].astype(np.float32)
            if i >= 23:
                q = q + np.float32(0.02 * (i // 23)) * rng.standard_normal(512).astype(np.float32)
            if i >= 56:
                q = rng.standard_normal(512).astype(np.float32) * 3.0
            res = gm.search(q, top_k=5)
            probes.append(q)
            exp_ids.append([r[0] for r in res])
            exp_scores.append([r[2] for r in res])
    np.savez_compressed(os.path.join(OUT, "search_cases.npz"), gallery_tag=np.array(tag), probes=np.array(probes, np.float32),
                        ids=np.array(exp_ids), scores=np.array(exp_scores, np.float64))

    # ---- recorded track result
    rec = json.load(open(f"{REF}/output/camera_captures/track_001/recognition_result.json"))
    keep = {k: rec[k] for k in ("recognized", "student_id", "name", "confidence", "method", "num_frames") if k in rec}
    keep["frame_matches"] = [{k: m[k] for k in ("frame", "student_id", "name", "score", "top_k_matches")} for m in rec["frame_matches"]]
    json.dump(keep, open(os.path.join(OUT, "track_001.json"), "w"), indent=1)

    # ---- preprocess arithmetic, reference lines executed verbatim
    import cv2
    pre = {}
    for i, S in enumerate([112, 224, 160, 112]):
        img = rng.integers(0, 256, (S, S, 3), dtype=np.uint8)
        img = cv2.GaussianBlur(img, (0, 0), 1.5)
        face_image = img
        if face_image.shape[:2] != (112, 112):                               # face_embedder.py:94-96
            face_image = cv2.resize(face_image, (112, 112), interpolation=cv2.INTER_LINEAR)
        bgr_img = face_image[:, :, ::-1]                                     # :99
        ada = ((bgr_img / 255.0 - 0.5) / 0.5).transpose(2, 0, 1).astype(np.float32)[None]   # :100-102
        arc = np.expand_dims(((bgr_img - 127.5) / 127.5).transpose(2, 0, 1), axis=0).astype(np.float32)  # :107-110
        pre[f"p{i}/img"] = img
        pre[f"p{i}/adaface"] = ada
        pre[f"p{i}/arcface"] = arc
    pre["num_cases"] = np.array(4)
    np.savez_compressed(os.path.join(OUT, "preprocess_cases.npz"), **pre)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
