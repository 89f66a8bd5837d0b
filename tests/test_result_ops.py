"""Result-side helpers after NMS (scale_boxes / clip_boxes / regularize_rboxes / scale_coords / clip_coords).

CPU part: the numpy oracle against the golden vectors of the LIVE reference (tests/golden/post/result_ops.npz, made by
oracle/make_golden_post.py) - bit-exact.  GPU part: the CUDA drop-ins (ultralytics_pro_b200.ops, through the C-ABI)
against the same golden vectors and against the oracle on batched rows - bit-exact (NaN == NaN)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro

PATH = os.path.join(os.path.dirname(__file__), "golden", "post", "result_ops.npz")


def _cases():
    z = np.load(PATH)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


Z, META = _cases()
IDS = [f"{i}-{m['kind']}" for i, m in enumerate(META)]


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def _rp(m):
    rp = m.get("ratio_pad")
    return None if rp is None else (tuple(rp[0]), tuple(rp[1]))


def _oracle(m, x):
    k = m["kind"]
    if k == "scale_boxes":
        return ro.scale_boxes_oracle(m["img1"], x, m["img0"], _rp(m), m["padding"], m["xywh"])
    if k == "clip_boxes":
        return ro.clip_boxes_oracle(x, m["shape"])
    if k == "scale_coords":
        return ro.scale_coords_oracle(m["img1"], x, m["img0"], _rp(m), m["normalize"], m["padding"])
    if k == "clip_coords":
        return ro.clip_coords_oracle(x, m["shape"])
    if k == "regularize_rboxes":
        return ro.regularize_rboxes_oracle(x)
    if k == "obb_result":
        return ro.obb_result_oracle(x, m["img1"], m["img0"])
    raise AssertionError(k)


@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_oracle_matches_reference_golden(i):
    got = _oracle(META[i], Z[f"c{i}_in"])
    assert _same(got, Z[f"c{i}_out"]), f"{META[i]}: max diff {np.nanmax(np.abs(got - Z[f'c{i}_out']))}"


def test_library_exports_scale_rows():
    import ctypes

    from ultralytics_pro_b200 import _cabi, build

    lib = ctypes.CDLL(build.build_library())
    assert hasattr(lib, "ypb_scale_rows") and "ypb_scale_rows" in _cabi.EXPORTS
    assert ctypes.sizeof(_cabi.ScaleXform) == 32


def test_cpu_tensor_raises():
    from ultralytics_pro_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.scale_boxes((640, 640), torch.zeros(3, 4), (480, 640))


def _ours(m, x, dev):
    from ultralytics_pro_b200 import ops

    t = torch.from_numpy(np.array(x, copy=True)).to(dev)
    k = m["kind"]
    if k == "scale_boxes":
        r = ops.scale_boxes(m["img1"], t, m["img0"], _rp(m), m["padding"], m["xywh"])
        assert r is t  # in place, like the reference
        return r
    if k == "clip_boxes":
        return ops.clip_boxes(t, m["shape"])
    if k == "scale_coords":
        return ops.scale_coords(m["img1"], t, m["img0"], _rp(m), m["normalize"], m["padding"])
    if k == "clip_coords":
        return ops.clip_coords(t, m["shape"])
    if k == "regularize_rboxes":
        r = ops.regularize_rboxes(t)
        assert r is not t
        return r
    if k == "obb_result":
        r = ops.regularize_rboxes(torch.cat([t[:, :4], t[:, -1:]], dim=-1))
        r[:, :4] = ops.scale_boxes(m["img1"], r[:, :4], m["img0"], xywh=True)  # strided (N, 4) view of (N, 5)
        return torch.cat([r, t[:, 4:6]], dim=-1)
    raise AssertionError(k)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_cuda_matches_reference_golden(cuda_device, i):
    got = _ours(META[i], Z[f"c{i}_in"], cuda_device).cpu().numpy()
    assert _same(got, Z[f"c{i}_out"]), f"{META[i]}: max diff {np.nanmax(np.abs(got - Z[f'c{i}_out']))}"


@pytest.mark.gpu
@pytest.mark.parametrize("rotated,kpt", [(False, None), (True, None), (False, (17, 3)), (False, (5, 2))])
def test_batched_scale_results_matches_oracle(cuda_device, rotated, kpt):
    """(B, max_det, cols) padded rows + device counts, per-image transforms, one launch; rows past the count untouched."""
    from ultralytics_pro_b200 import ops

    rng = np.random.default_rng(7)
    B, M = 5, 40
    extra = 1 if rotated else (kpt[0] * kpt[1] if kpt else 0)
    rows = rng.uniform(-20, 700, size=(B, M, 6 + extra)).astype(np.float32)
    if rotated:
        rows[..., 6] = rng.uniform(-0.8, 2.4, size=(B, M))
    counts = np.array([0, 1, 17, 40, 33], dtype=np.int32)
    img1 = (640, 640)
    shapes = [(480, 640, 3), (1080, 1920), (640, 640), (333, 500, 3), (2000, 1500)]
    t = torch.from_numpy(rows.copy()).to(cuda_device)
    c = torch.from_numpy(counts).to(cuda_device)
    ops.scale_results(t, c, img1, shapes, rotated=rotated, kpt_shape=kpt)
    got = t.cpu().numpy()
    for b in range(B):
        n = counts[b]
        want = rows[b].copy()
        if rotated:
            r = ro.regularize_rboxes_oracle(np.concatenate([want[:n, :4], want[:n, 6:7]], -1))
            r[:, :4] = ro.scale_boxes_oracle(img1, r[:, :4], shapes[b], xywh=True)
            want[:n, :4], want[:n, 6] = r[:, :4], r[:, 4]
        else:
            want[:n, :4] = ro.scale_boxes_oracle(img1, want[:n, :4], shapes[b])
            if kpt:
                k = want[:n, 6:6 + kpt[0] * kpt[1]].reshape(n, kpt[0], kpt[1])
                want[:n, 6:6 + kpt[0] * kpt[1]] = ro.scale_coords_oracle(img1, k, shapes[b]).reshape(n, kpt[0] * kpt[1])
        assert _same(got[b], want), f"image {b}"


@pytest.mark.gpu
def test_scale_boxes_on_nms_rows_view(cuda_device):
    """detect/predict.py:120: `pred[:, :4] = ops.scale_boxes(img.shape[2:], pred[:, :4], orig_img.shape)` on a (n, 6) row view."""
    from ultralytics_pro_b200 import ops

    rng = np.random.default_rng(11)
    pred = rng.uniform(-5, 650, size=(50, 6)).astype(np.float32)
    t = torch.from_numpy(pred.copy()).to(cuda_device)
    t[:, :4] = ops.scale_boxes((640, 640), t[:, :4], (720, 1280, 3))
    want = pred.copy()
    want[:, :4] = ro.scale_boxes_oracle((640, 640), pred[:, :4], (720, 1280, 3))
    assert _same(t.cpu().numpy(), want)
    empty = torch.zeros((0, 6), device=cuda_device)
    assert ops.scale_boxes((640, 640), empty[:, :4], (720, 1280)).shape == (0, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("rotated", [False, True])
def test_fused_gather_rescale_equals_nms_then_scale(cuda_device, rotated):
    """postprocess_from_head(orig_shapes=...) == postprocess_from_head(...) followed by the oracle's construct_result scaling."""
    from tests.helpers import small_cfg
    from ultralytics_pro_b200.head import postprocess_from_head
    from ultralytics_pro_b200.pipeline import HeadPostProcessor
    from ultralytics_pro_b200.synth import make_head_batch

    cfg = small_cfg("scale", imgsz=160, nc=15 if rotated else 80, batch=3, rotated=rotated, objects=6)
    levels, ang = make_head_batch(cfg, batch=3, seed=5)
    dl = [lv.to(cuda_device) for lv in levels]
    da = ang.to(cuda_device) if ang is not None else None
    shapes = [(120, 160, 3), (480, 640), (333, 250, 3)]
    img = (160, 160)
    plain = postprocess_from_head(dl, cfg.strides, cfg.nc, 0.25, 0.7, angle_logits=da)
    fused = postprocess_from_head(dl, cfg.strides, cfg.nc, 0.25, 0.7, angle_logits=da, img_shape=img, orig_shapes=shapes)
    pp = HeadPostProcessor(cfg.nc, cfg.strides, 0.25, 0.7, rotated=rotated, scale_to_original=True)
    pp.set_image_shapes(dl, img, shapes)
    piped = pp(dl, da)
    assert sum(p.shape[0] for p in plain) > 0
    for b in range(3):
        want = plain[b].cpu().numpy().copy()
        if rotated:
            r = ro.regularize_rboxes_oracle(np.concatenate([want[:, :4], want[:, 6:7]], -1))
            r[:, :4] = ro.scale_boxes_oracle(img, r[:, :4], shapes[b], xywh=True)
            want[:, :4], want[:, 6] = r[:, :4], r[:, 4]
        else:
            want[:, :4] = ro.scale_boxes_oracle(img, want[:, :4], shapes[b])
        assert _same(fused[b].cpu().numpy(), want), f"image {b}"
        assert _same(piped[b].cpu().numpy(), want), f"image {b} (HeadPostProcessor)"
