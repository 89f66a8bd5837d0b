"""Generate tests/golden/*.npz by running the LIVE reference (Chriz122/ultralytics_pro from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the reference tree is not on the GPU box):

    python oracle/make_golden.py

Every fixture stores the seeded synthetic head tensors (small grids so the files stay small), the reference's
`Detect._inference` / `OBB.forward` output, and the reference's `non_max_suppression(..., return_idxs=True)` rows and
kept anchor indices for a set of argument combinations.  `max_time_img` is raised so the wall-clock guard
(nms.py:81,162-164) never fires; the input is cloned per call because the reference rewrites it in place (nms.py:86).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from tests.helpers import make_scores_unique  # noqa: E402
from ultralytics_pro_b200.synth import HeadConfig, make_head_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, config, batch, seed, list of NMS kwargs
    ("detect_160", HeadConfig("detect_160", 160, (8, 16, 32), 80, 2, objects=6), 2, 101, [
        dict(conf_thres=0.25, iou_thres=0.7),
        dict(conf_thres=0.25, iou_thres=0.45, agnostic=True),
        dict(conf_thres=0.05, iou_thres=0.6, classes=[0, 5, 17, 33, 79]),
        dict(conf_thres=0.25, iou_thres=0.7, max_det=5),
        dict(conf_thres=0.001, iou_thres=0.7, multi_label=True),
        dict(conf_thres=0.001, iou_thres=0.7, multi_label=True, max_nms=300),
    ]),
    ("detect_p6_256", HeadConfig("detect_p6_256", 256, (8, 16, 32, 64), 20, 2, objects=8), 2, 202, [
        dict(conf_thres=0.25, iou_thres=0.7),
        dict(conf_thres=0.1, iou_thres=0.5, agnostic=True),
    ]),
    ("obb_192", HeadConfig("obb_192", 192, (8, 16, 32), 15, 2, rotated=True, objects=8), 2, 303, [
        dict(conf_thres=0.25, iou_thres=0.7, rotated=True),
        dict(conf_thres=0.01, iou_thres=0.3, rotated=True),
    ]),
]


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, cfg, batch, seed, calls in CASES:
        levels, ang = make_head_batch(cfg, batch=batch, seed=seed)
        ch = tuple(16 * (2 ** i) for i in range(len(cfg.strides)))
        if cfg.rotated:
            head = ref.OBB(nc=cfg.nc, ne=1, ch=ch)
        else:
            head = ref.Detect(nc=cfg.nc, ch=ch)
        head.stride = torch.tensor([float(s) for s in cfg.strides])
        head.eval()
        with torch.inference_mode():
            if cfg.rotated:
                import math

                head.angle = (ang.sigmoid() - 0.25) * math.pi  # head.py:1031
                y = head._inference([lv.clone() for lv in levels])
                y = torch.cat([y, head.angle], 1)  # head.py:1038
            else:
                y = head._inference([lv.clone() for lv in levels])
        y = y.clone()
        raw = y.clone()
        # tie-free scores above the smallest conf used, so that the reference's unstable argsort cannot reorder rows
        y = make_scores_unique(y, cfg.nc, min(c["conf_thres"] for c in calls))
        blob = {f"level{i}": lv.numpy() for i, lv in enumerate(levels)}
        if ang is not None:
            blob["angle_logits"] = ang.numpy()
        blob["decoded_raw"] = raw.numpy()  # the reference's decode, untouched
        blob["decoded"] = y.numpy()        # the tensor fed to the reference's NMS (scores made tie-free)
        meta = {"name": name, "imgsz": cfg.imgsz, "strides": list(cfg.strides), "nc": cfg.nc, "reg_max": cfg.reg_max,
                "rotated": cfg.rotated, "batch": batch, "seed": seed, "calls": []}
        for ci, kw in enumerate(calls):
            with torch.inference_mode():
                out, keep = ref.non_max_suppression(y.clone(), nc=cfg.nc, max_time_img=1e9, return_idxs=True, **kw)
            for b in range(batch):
                blob[f"call{ci}_rows{b}"] = out[b].numpy().astype(np.float32)
                blob[f"call{ci}_idx{b}"] = keep[b].reshape(-1).numpy().astype(np.int64)
            meta["calls"].append(kw)
            print(name, kw, [int(o.shape[0]) for o in out])
        blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
