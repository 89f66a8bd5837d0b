"""Generate tests/golden/post/result_ops.npz by running the LIVE reference's result-side helpers (utils/ops.py).

TEST INFRASTRUCTURE.  Build container only:   python oracle/make_golden_post.py
Each case stores the seeded input, the call's arguments (meta JSON) and the reference's output.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "post", "result_ops.npz")

SHAPES = [  # (network input HW, original image HW[C], ratio_pad)
    ((640, 640), (480, 640, 3), None),
    ((384, 640), (1080, 1920, 3), None),
    ((640, 640), (427, 640), None),
    ((640, 480), (1333, 999, 3), None),
    ((1024, 1024), (3000, 4000, 3), None),
    ((640, 640), (720, 1280), ((0.5, 0.5), (0.0, 140.0))),
    ((640, 640), (500, 375), ((1.28, 1.28), (80.5, 0.25))),
]


def boxes_input(rng, n, h, w):
    xy = rng.uniform(-40, max(h, w) + 40, size=(n, 2))
    wh = rng.uniform(1, max(h, w) * 0.6, size=(n, 2))
    b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    b[0] = [0, 0, w, h]
    b[1] = [-0.0, 1e-3, w + 1e-3, h - 1e-3]
    if n > 4:
        b[2, 1] = np.nan
        b[3] = [np.inf, -np.inf, 1e9, -1e9]
    return b


def main():
    ref = load_reference()
    ops = ref.ops
    rng = np.random.default_rng(20261017)
    blob, meta = {}, []

    def add(kind, args, inp, out):
        i = len(meta)
        blob[f"c{i}_in"] = np.asarray(inp, dtype=np.float32)
        blob[f"c{i}_out"] = np.asarray(out, dtype=np.float32)
        meta.append({"kind": kind, **args})
        print(i, kind, args, blob[f"c{i}_in"].shape)

    for s1, s0, rp in SHAPES:
        for padding in (True, False):
            for xywh in (False, True):
                b = boxes_input(rng, 41, *s1)
                out = ops.scale_boxes(s1, torch.from_numpy(b.copy()), s0, ratio_pad=rp, padding=padding, xywh=xywh)
                add("scale_boxes", dict(img1=list(s1), img0=list(s0), ratio_pad=rp, padding=padding, xywh=xywh), b, out.numpy())
        b = boxes_input(rng, 33, *s1)
        add("clip_boxes", dict(shape=list(s0)), b, ops.clip_boxes(torch.from_numpy(b.copy()), s0).numpy())
        for normalize in (False, True):
            k = rng.uniform(-30, max(s1) + 30, size=(9, 17, 3)).astype(np.float32)
            k[..., 2] = rng.uniform(0, 1, size=(9, 17))
            k[0, 0, 0] = np.nan
            out = ops.scale_coords(s1, torch.from_numpy(k.copy()), s0, ratio_pad=rp, normalize=normalize)
            add("scale_coords", dict(img1=list(s1), img0=list(s0), ratio_pad=rp, normalize=normalize, padding=True), k, out.numpy())
        k = rng.uniform(-30, max(s1) + 30, size=(64, 2)).astype(np.float32)
        out = ops.scale_coords(s1, torch.from_numpy(k.copy()), s0, ratio_pad=rp, padding=False)
        add("scale_coords", dict(img1=list(s1), img0=list(s0), ratio_pad=rp, normalize=False, padding=False), k, out.numpy())
        k = rng.uniform(-30, max(s0[:2]) + 30, size=(5, 17, 2)).astype(np.float32)
        add("clip_coords", dict(shape=list(s0)), k, ops.clip_coords(torch.from_numpy(k.copy()), s0).numpy())

    # rotated boxes: the OBB head's angle range [-pi/4, 3pi/4) (head.py:1031) plus the branch points of ops.py:631-634
    n = 200
    rb = np.concatenate([rng.uniform(0, 1024, (n, 2)), rng.uniform(2, 400, (n, 2)),
                         rng.uniform(-math.pi / 4, 3 * math.pi / 4, (n, 1))], 1).astype(np.float32)
    edge = [0.0, -0.0, math.pi / 2, np.float32(math.pi / 2), np.nextafter(np.float32(math.pi / 2), np.float32(0)),
            math.pi, np.float32(math.pi), -math.pi / 4, 3 * math.pi / 4, -1e-7, 1e-7, 2.0, -2.0, 7.0, -7.0, 1.5707964]
    rb[: len(edge), 4] = np.asarray(edge, dtype=np.float32)
    add("regularize_rboxes", {}, rb, ops.regularize_rboxes(torch.from_numpy(rb.copy())).numpy())
    for s1, s0, _ in SHAPES[:5]:
        pred = np.concatenate([rb[:, :4], rng.uniform(0.25, 1, (n, 1)), rng.integers(0, 15, (n, 1)), rb[:, 4:5]], 1).astype(np.float32)
        p = torch.from_numpy(pred.copy())
        # models/yolo/obb/predict.py:59-61, executed with the reference's own functions
        r = ops.regularize_rboxes(torch.cat([p[:, :4], p[:, -1:]], dim=-1))
        r[:, :4] = ops.scale_boxes(s1, r[:, :4], s0, xywh=True)
        add("obb_result", dict(img1=list(s1), img0=list(s0)), pred, torch.cat([r, p[:, 4:6]], dim=-1).numpy())

    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    make_kpts(ref)
    make_masks(ref)
    make_matching(ref)
    make_nms_model(ref)
    make_topk(ref)


def make_topk(ref):
    """Detect.postprocess of the live reference (nn/modules/head.py:193-214) on tie-free sigmoid scores."""
    out_path = os.path.join(ROOT, "tests", "golden", "post", "topk.npz")
    g = torch.Generator().manual_seed(31)
    blob, meta = {}, []
    for name, b, a, nc, max_det in [("v10_small", 2, 525, 80, 300), ("few_anchors", 3, 40, 5, 300), ("nc1", 2, 300, 1, 50)]:
        boxes = torch.rand(b, a, 4, generator=g) * 160
        scores = torch.sigmoid(torch.randn(b, a, nc, generator=g) * 2 - 3)
        flat = scores.reshape(b, -1)
        for i in range(b):  # make every score of an image distinct
            srt, idx = torch.sort(flat[i])
            for _ in range(4):
                dup = torch.zeros_like(srt, dtype=torch.bool)
                dup[1:] = srt[1:] <= srt[:-1]
                if not dup.any():
                    break
                srt = torch.where(dup, torch.nextafter(torch.roll(srt, 1), torch.full_like(srt, 2.0)), srt)
                srt = torch.cummax(srt, 0).values
            flat[i, idx] = srt
        preds = torch.cat([boxes, flat.reshape(b, a, nc)], -1)
        out = ref.Detect.postprocess(preds.clone(), max_det, nc)
        i = len(meta)
        blob[f"t{i}_preds"], blob[f"t{i}_out"] = preds.numpy(), out.numpy()
        meta.append(dict(name=name, max_det=max_det, nc=nc))
        print(name, tuple(out.shape))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


def make_nms_model(ref):
    """NMSModel.forward of the live reference (engine/exporter.py:1389-1481) fed by a stub model that returns a decoded tensor."""
    import types

    from ultralytics.engine.exporter import NMSModel

    from tests.helpers import make_scores_unique
    from oracle.postproc_oracle import decode_oracle
    from ultralytics_pro_b200.synth import HeadConfig, make_head_batch

    out_path = os.path.join(ROOT, "tests", "golden", "post", "nms_model.npz")
    blob, meta = {}, []
    for name, imgsz, nc, extra, kw in [("detect", 160, 80, 0, dict(conf=0.25, iou=0.45, max_det=20, agnostic_nms=False)),
                                       ("detect_agnostic", 160, 80, 0, dict(conf=0.1, iou=0.6, max_det=300, agnostic_nms=True)),
                                       ("segment_extras", 128, 20, 32, dict(conf=0.25, iou=0.7, max_det=50, agnostic_nms=False))]:
        cfg = HeadConfig(name, imgsz, (8, 16, 32), nc, 2, objects=7)
        levels, _ = make_head_batch(cfg, batch=2, seed=len(meta) + 11)
        y = decode_oracle(levels, cfg.strides, nc, xyxy=True)  # the export decode: corners (head.py:189)
        y = make_scores_unique(y, nc, kw["conf"])
        if extra:
            y = torch.cat([y, torch.randn(2, extra, y.shape[2], generator=torch.Generator().manual_seed(5))], 1)

        class Stub(torch.nn.Module):
            task = "detect"
            names = {i: str(i) for i in range(nc)}

            def forward(self, x):
                return y.clone()

        args = types.SimpleNamespace(format="torchscript", dynamic=False, batch=2, opset=None, int8=False, **kw)
        with torch.inference_mode():
            out = NMSModel(Stub(), args)(torch.zeros(2, 3, imgsz, imgsz))
        i = len(meta)
        blob[f"n{i}_pred"], blob[f"n{i}_out"] = y.numpy(), out.numpy()
        meta.append(dict(name=name, imgsz=imgsz, nc=nc, extra=extra, **kw))
        print(name, tuple(out.shape), int((out[..., 4] > 0).sum()))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


def make_matching(ref):
    """BaseValidator.match_predictions + box_iou of the live reference (engine/validator.py:267-307, utils/metrics.py:54)."""
    import types

    from ultralytics.engine.validator import BaseValidator

    out_path = os.path.join(ROOT, "tests", "golden", "post", "matching.npz")
    iouv = torch.linspace(0.5, 0.95, 10)
    ns = types.SimpleNamespace(iouv=iouv)
    g = torch.Generator().manual_seed(99)
    blob, meta = {}, []
    for name, m, n, nc, jitter in [("few", 6, 20, 3, 3.0), ("crowd", 120, 300, 5, 6.0), ("one_class", 40, 90, 1, 10.0),
                                   ("no_match", 10, 30, 4, 400.0), ("many_labels", 700, 300, 2, 4.0)]:
        xy = torch.rand(m, 2, generator=g) * 500
        wh = torch.rand(m, 2, generator=g) * 120 + 8
        gt = torch.cat([xy, xy + wh], 1)
        gcls = torch.randint(0, nc, (m,), generator=g).float()
        src = torch.randint(0, m, (n,), generator=g)
        pred = gt[src] + torch.randn(n, 4, generator=g) * jitter
        pcls = torch.where(torch.rand(n, generator=g) < 0.8, gcls[src], torch.randint(0, nc, (n,), generator=g).float())
        iou = ref.metrics.box_iou(gt, pred)
        tp = BaseValidator.match_predictions(ns, pcls, gcls, iou)
        i = len(meta)
        blob[f"q{i}_gt"], blob[f"q{i}_gcls"], blob[f"q{i}_pred"], blob[f"q{i}_pcls"] = gt.numpy(), gcls.numpy(), pred.numpy(), pcls.numpy()
        blob[f"q{i}_iou"], blob[f"q{i}_tp"] = iou.numpy(), tp.numpy()
        meta.append(dict(name=name, m=m, n=n))
        print(name, tuple(tp.shape), int(tp.sum()))
    blob["iouv"] = iouv.numpy()
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


def make_masks(ref):
    """process_mask / process_mask_native (ops.py:489-541) of the live reference; n >= 50 so that the CPU run takes the
    same crop_mask branch as CUDA tensors do (ops.py:482-486)."""
    out_path = os.path.join(ROOT, "tests", "golden", "post", "masks.npz")
    ops = ref.ops
    g = torch.Generator().manual_seed(4242)
    blob, meta = {}, []
    for name, (mh, mw), shape, n, kind, dtype in [
        ("up_4x", (40, 40), (160, 160), 56, "process_mask_up", torch.float32),
        ("low_res", (40, 40), (160, 160), 56, "process_mask", torch.float32),
        ("up_rect_f16", (24, 40), (96, 160), 50, "process_mask_up", torch.float16),
        ("native_letterbox", (40, 40), (120, 213), 52, "process_mask_native", torch.float32),
        ("native_tall", (40, 40), (333, 250), 50, "process_mask_native", torch.float32),
    ]:
        # smooth prototypes (sums of a few low-frequency waves) so that masks look like blobs, plus noise
        yy, xx = torch.meshgrid(torch.linspace(0, 1, mh), torch.linspace(0, 1, mw), indexing="ij")
        protos = torch.stack([torch.sin(6.28 * (torch.rand(1, generator=g) * 3 * xx + torch.rand(1, generator=g) * 3 * yy
                                                + torch.rand(1, generator=g))) for _ in range(32)])
        protos = (protos + 0.05 * torch.randn(32, mh, mw, generator=g)).to(dtype)
        coef = torch.randn(n, 32, generator=g)
        H, W = (mh * 4, mw * 4) if kind != "process_mask_native" else shape
        xy = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.8, H * 0.8])
        wh = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.5, H * 0.5]) + 2
        boxes = torch.cat([xy, xy + wh], 1)
        boxes[0] = torch.tensor([0.0, 0.0, float(W), float(H)])
        boxes[1] = torch.tensor([8.0, 12.0, 8.0, 40.0])       # empty in x
        boxes[2] = torch.tensor([-30.0, -30.0, W + 30.0, 16.0])
        boxes[3] = torch.tensor([10.5, 10.5, 11.5, 11.5])     # inside one prototype cell
        if kind == "process_mask_up":
            out = ops.process_mask(protos, coef.clone(), boxes.clone(), shape, upsample=True)
        elif kind == "process_mask":
            out = ops.process_mask(protos, coef.clone(), boxes.clone(), shape, upsample=False)
        else:
            out = ops.process_mask_native(protos, coef.clone(), boxes.clone(), shape)
        i = len(meta)
        blob[f"m{i}_protos"] = protos.float().numpy()
        blob[f"m{i}_coef"] = coef.numpy()
        blob[f"m{i}_boxes"] = boxes.numpy()
        blob[f"m{i}_out"] = np.packbits(out.numpy().astype(bool))
        meta.append(dict(name=name, kind=kind, shape=list(shape), out_shape=list(out.shape), dtype=str(dtype).split(".")[1]))
        print(name, tuple(out.shape), int(out.sum()))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


def make_kpts(ref):
    """Pose.kpts_decode (head.py:1254-1273) of the live reference on seeded raw keypoint logits."""
    out_path = os.path.join(ROOT, "tests", "golden", "post", "kpts.npz")
    blob, meta = {}, []
    g = torch.Generator().manual_seed(77)
    for name, imgsz, strides, kpt_shape, dtype in [
        ("pose17x3_f32", 96, (8, 16, 32), (17, 3), torch.float32),
        ("pose5x2_f32", 96, (8, 16, 32), (5, 2), torch.float32),
        ("pose17x3_p6_f32", 128, (8, 16, 32, 64), (17, 3), torch.float32),
        ("pose17x3_bf16", 96, (8, 16, 32), (17, 3), torch.bfloat16),
        ("pose17x3_wide_f16", 1280, (8,), (3, 3), torch.float16),  # 160-wide grid: half-integer anchors round in 16 bit
    ]:
        hw = [(imgsz // s, imgsz // s) for s in strides]
        if name.endswith("wide_f16"):
            hw = [(2, 160)]
        a = sum(h * w for h, w in hw)
        head = ref.head.Pose(nc=1, kpt_shape=kpt_shape, ch=tuple(16 for _ in strides))
        head.stride = torch.tensor([float(s) for s in strides])
        feats = [torch.zeros(1, 16, h, w, dtype=dtype) for h, w in hw]
        anc, st = ref.tal.make_anchors(feats, head.stride, 0.5)
        head.anchors, head.strides = anc.transpose(0, 1), st.transpose(0, 1)  # head.py:163
        head.export = False
        kp = (torch.randn(2, kpt_shape[0] * kpt_shape[1], a, generator=g) * 2.5).to(dtype)
        with torch.inference_mode():
            out = head.kpts_decode(2, kp)
        i = len(meta)
        blob[f"k{i}_in"] = kp.float().numpy()
        blob[f"k{i}_out"] = out.float().numpy()
        meta.append(dict(name=name, level_hw=hw, strides=list(strides), kpt_shape=list(kpt_shape), dtype=str(dtype).split(".")[1]))
        print(name, tuple(kp.shape))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
