"""Segment / Pose riders (part 1): Pose.kpts_decode (head.py:1254-1273) as a dense kernel and as a
per-kept-anchor decode inside the fused path; Segment mask coefficients (head.py:831,837) gathered for kept rows.

CPU: the oracle restatement against the live-reference golden vectors (tests/golden/post/kpts.npz) - bit-exact.
GPU: kernels against the golden vectors (x, y bit-exact in every dtype, visibility = sigmoid within 1e-5 / 1e-2), the
fused path against decode + cat + non_max_suppression (bit-exact) and against the oracle chain."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro
from oracle.postproc_oracle import decode_oracle, nms_oracle
from tests.helpers import make_scores_unique, small_cfg
from ultralytics_pro_b200.synth import make_head_batch

PATH = os.path.join(os.path.dirname(__file__), "golden", "post", "kpts.npz")
Z = np.load(PATH)
META = json.loads(bytes(Z["meta"]).decode())
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


@pytest.mark.parametrize("i", range(len(META)), ids=[m["name"] for m in META])
def test_oracle_kpts_matches_reference_golden(i):
    m = META[i]
    kp = torch.from_numpy(Z[f"k{i}_in"]).to(DT[m["dtype"]])
    got = ro.kpts_decode_oracle(kp, m["level_hw"], m["strides"], m["kpt_shape"]).float().numpy()
    assert np.array_equal(got, Z[f"k{i}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=[m["name"] for m in META])
def test_cuda_kpts_decode_matches_reference_golden(cuda_device, i):
    from ultralytics_pro_b200.head import decode_keypoints

    m = META[i]
    nk, ndim = m["kpt_shape"]
    kp = torch.from_numpy(Z[f"k{i}_in"]).to(DT[m["dtype"]]).to(cuda_device)
    got = decode_keypoints(kp, m["level_hw"], m["strides"], m["kpt_shape"])
    assert got.dtype == kp.dtype and got.shape == kp.shape
    got = got.float().cpu().numpy().reshape(2, nk, ndim, -1)
    want = Z[f"k{i}_out"].reshape(2, nk, ndim, -1)
    assert np.array_equal(got[:, :, :2], want[:, :, :2]), "x / y must be bit-exact (exact fp arithmetic, per-op rounding)"
    if ndim == 3:
        tol = 1e-5 if m["dtype"] == "float32" else 1e-2
        rel = np.abs(got[:, :, 2] - want[:, :, 2]) / np.maximum(np.abs(want[:, :, 2]), 1e-30)
        assert rel.max() <= tol, f"visibility sigmoid rel err {rel.max()}"


def _pose_inputs(dev, dtype=torch.float32, batch=3, seed=21, kpt_shape=(17, 3)):
    cfg = small_cfg("pose", imgsz=160, nc=1, batch=batch, objects=7)
    levels, _ = make_head_batch(cfg, batch=batch, seed=seed, dtype=dtype)
    g = torch.Generator().manual_seed(seed)
    kp = (torch.randn(batch, kpt_shape[0] * kpt_shape[1], cfg.anchors, generator=g) * 2.0).to(dtype)
    return cfg, levels, kp


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_pose_riders_equal_decode_cat_nms(cuda_device, dtype):
    from ultralytics_pro_b200.head import decode_head, decode_keypoints, postprocess_from_head
    from ultralytics_pro_b200.nms import non_max_suppression

    kpt_shape = (17, 3)
    cfg, levels, kp = _pose_inputs(cuda_device, dtype)
    dl, dk = [lv.to(cuda_device) for lv in levels], kp.to(cuda_device)
    dense = torch.cat([decode_head(dl, cfg.strides, cfg.nc), decode_keypoints(dk, cfg.level_hw, cfg.strides, kpt_shape)], 1)  # head.py:1252
    two, two_idx = non_max_suppression(dense, 0.25, 0.7, nc=1, return_idxs=True)
    fused, fused_idx = postprocess_from_head(dl, cfg.strides, 1, 0.25, 0.7, kpt_logits=dk, kpt_shape=kpt_shape, return_idxs=True)
    assert sum(t.shape[0] for t in two) > 0
    for b in range(len(two)):
        assert fused[b].shape == two[b].shape and fused[b].shape[1] == 6 + 51
        assert torch.equal(fused_idx[b], two_idx[b])
        assert torch.equal(fused[b], two[b]), f"image {b}: fused rider rows differ from decode+cat+NMS"


@pytest.mark.gpu
def test_fused_pose_riders_against_oracle_chain_and_scaling(cuda_device):
    from ultralytics_pro_b200.head import postprocess_from_head

    kpt_shape = (17, 3)
    cfg, levels, kp = _pose_inputs(cuda_device)
    y = torch.cat([decode_oracle(levels, cfg.strides, cfg.nc), ro.kpts_decode_oracle(kp, cfg.level_hw, cfg.strides, kpt_shape)], 1)
    want, want_idx = nms_oracle(y, 0.25, 0.7, nc=1)
    dl, dk = [lv.to(cuda_device) for lv in levels], kp.to(cuda_device)
    got, got_idx = postprocess_from_head(dl, cfg.strides, 1, 0.25, 0.7, kpt_logits=dk, kpt_shape=kpt_shape, return_idxs=True)
    shapes = [(120, 160, 3), (480, 640), (333, 250, 3)]
    scaled = postprocess_from_head(dl, cfg.strides, 1, 0.25, 0.7, kpt_logits=dk, kpt_shape=kpt_shape, img_shape=(160, 160), orig_shapes=shapes)
    for b in range(len(want)):
        assert torch.equal(got_idx[b].cpu(), want_idx[b].view(-1)), f"image {b}: kept anchors differ from the oracle"
        g, w = got[b].cpu(), want[b]
        assert torch.allclose(g[:, :6], w[:, :6], rtol=1e-5, atol=1e-5 * cfg.imgsz)
        gk, wk = g[:, 6:].view(-1, 17, 3), w[:, 6:].view(-1, 17, 3)
        assert torch.equal(gk[..., :2], wk[..., :2]), "keypoint x / y bit-exact"
        assert torch.allclose(gk[..., 2], wk[..., 2], rtol=1e-5, atol=0)
        # construct_result scaling (detect/predict.py:120 + pose/predict.py:73-75) folded into the gather
        exp = g.numpy().copy()
        exp[:, :4] = ro.scale_boxes_oracle((160, 160), exp[:, :4], shapes[b])
        exp[:, 6:] = ro.scale_coords_oracle((160, 160), exp[:, 6:].reshape(-1, 17, 3), shapes[b]).reshape(len(exp), -1)
        assert np.array_equal(scaled[b].cpu().numpy(), exp), f"image {b}: fused rescale of boxes / keypoints"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_fused_segment_riders_equal_decode_cat_nms(cuda_device, dtype):
    """Segment: (B, 32, A) mask coefficients ride as extras (head.py:837) - fused gather == dense path, bit for bit."""
    from ultralytics_pro_b200.head import decode_head, postprocess_from_head
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = small_cfg("seg", imgsz=160, nc=80, batch=2, objects=7)
    levels, _ = make_head_batch(cfg, batch=2, seed=31, dtype=dtype)
    mc = torch.randn(2, 32, cfg.anchors, generator=torch.Generator().manual_seed(3)).to(dtype)
    dl, dm = [lv.to(cuda_device) for lv in levels], mc.to(cuda_device)
    dense = torch.cat([decode_head(dl, cfg.strides, cfg.nc), dm], 1)
    two = non_max_suppression(dense, 0.25, 0.7, nc=cfg.nc)
    fused = postprocess_from_head(dl, cfg.strides, cfg.nc, 0.25, 0.7, mask_coeffs=dm)
    assert sum(t.shape[0] for t in two) > 0
    for b in range(2):
        assert fused[b].shape[1] == 38 and torch.equal(fused[b], two[b])
    # non-contiguous coefficient tensor (a channel slice of a wider tensor)
    wide = torch.randn(2, 40, cfg.anchors, device=cuda_device).to(dtype)
    sl = wide[:, 4:36]
    f2 = postprocess_from_head(dl, cfg.strides, cfg.nc, 0.25, 0.7, mask_coeffs=sl)
    t2 = non_max_suppression(torch.cat([decode_head(dl, cfg.strides, cfg.nc), sl], 1), 0.25, 0.7, nc=cfg.nc)
    for b in range(2):
        assert torch.equal(f2[b], t2[b])


def test_rider_argument_errors():
    from ultralytics_pro_b200 import _cabi

    assert {"ypb_nms_from_head_riders", "ypb_kpts_decode"} <= set(_cabi.EXPORTS)
