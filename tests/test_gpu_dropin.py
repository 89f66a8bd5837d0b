"""GPU tests of the BOUND drop-ins: what ``patch.install()`` rebinds is exercised here on CUDA
tensors through the very wrappers it installs, on stub modules that carry the attributes the reference sets
(head.py:70-93 Detect, :1026-1042 OBB, :1254-1273 Pose, validator.py:148 iouv) - the reference package itself cannot
travel to the GPU box.  Every kernel branch is taken with a reference function that raises, so a silent fall-through to
"the reference" cannot pass.  The steady-state calls run under ``torch.cuda.set_sync_debug_mode("error")``: the drop-in
``_inference`` / ``decode_bboxes`` / ``DFL.forward`` / ``kpts_decode`` issue no device->host synchronisation once their
per-module host caches are warm (the reference's ``_inference`` has none either on a shape-cache hit)."""
import contextlib
import math
import types

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro
from oracle.postproc_oracle import anchor_table, decode_oracle, dfl_expect, nms_oracle, obb_forward_oracle
from tests.helpers import assert_rows_equal, small_cfg
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

pytestmark = pytest.mark.gpu


def _boom(*a, **k):
    raise AssertionError("the saved reference function was called: the CUDA branch of the wrapper was not taken")


@contextlib.contextmanager
def no_host_sync():
    torch.cuda.set_sync_debug_mode("error")
    try:
        yield
    finally:
        torch.cuda.set_sync_debug_mode("default")


class StubDetect:
    """Attribute surface of ``Detect`` (head.py:70-93) without the convolutions."""

    dynamic = False
    export = False
    format = None
    end2end = False
    max_det = 300
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    legacy = False
    xyxy = False
    training = False

    def __init__(self, nc, strides, device, reg_max=16):
        self.nc, self.nl, self.reg_max = nc, len(strides), reg_max
        self.no = nc + 4 * reg_max
        self.stride = torch.tensor([float(s) for s in strides], device=device)  # BaseModel._apply moved it (tasks.py:1193-1210)


def _bind(obj, name, wrapper):
    setattr(obj, name, types.MethodType(wrapper, obj))
    return obj


def _detect(cfg, dev, **attrs):
    from ultralytics_pro_b200 import head, patch

    m = StubDetect(cfg.nc, cfg.strides, dev)
    for k, v in attrs.items():
        setattr(m, k, v)
    _bind(m, "_inference", patch._wrap_inference(_boom, head.detect_inference))
    _bind(m, "decode_bboxes", patch._wrap_decode_bboxes(_boom, head.detect_decode_bboxes))
    return m


@pytest.fixture(autouse=True)
def _lazy_off():
    from ultralytics_pro_b200 import lazy

    lazy.ENABLED = False
    yield
    lazy.ENABLED = False


# ------------------------------------------------------------------------------------------------------------------
# Detect._inference
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["plain", "xyxy", "end2end", "dynamic"])
def test_bound_inference_matches_decode_and_caches_like_the_reference(cuda_device, variant):
    from ultralytics_pro_b200.head import decode_head

    cfg = small_cfg(batch=3)
    levels, _ = make_head_batch(cfg, seed=31)
    dl = [lv.to(cuda_device) for lv in levels]
    m = _detect(cfg, cuda_device, xyxy=variant == "xyxy", end2end=variant == "end2end", dynamic=variant == "dynamic")
    xyxy = variant in ("xyxy", "end2end")
    with torch.inference_mode():
        y = m._inference(dl)  # first call: builds the host caches (one read-back of self.stride, like make_anchors)
        with no_host_sync():
            y2 = m._inference(dl)
    assert type(y) is torch.Tensor and y.dtype == dl[0].dtype and y.shape == (3, 4 + cfg.nc, cfg.anchors)
    assert torch.equal(y, y2) and torch.equal(y, decode_head(dl, cfg.strides, cfg.nc, xyxy=xyxy))
    want = decode_oracle(levels, cfg.strides, cfg.nc, xyxy=xyxy)
    assert float((y.cpu() - want).abs().max()) <= 1e-5 * cfg.imgsz
    # head.py:163-165: the module keeps caching anchors (2, A), strides (1, A) and the input shape
    anc, srow = anchor_table(cfg.level_hw, cfg.strides)
    assert m.shape == dl[0].shape
    assert torch.equal(m.anchors.cpu(), anc) and m.anchors.shape == (2, cfg.anchors)
    assert torch.equal(m.strides.cpu(), srow) and m.strides.shape == (1, cfg.anchors)


def test_bound_inference_half_and_obb(cuda_device):
    from ultralytics_pro_b200 import head, patch

    cfg = CONFIGS["c5_obb_1024_b16"]
    levels, ang = make_head_batch(cfg, batch=2, seed=33)
    dl, da = [lv.to(cuda_device) for lv in levels], ang.to(cuda_device)
    m = _detect(cfg, cuda_device)
    m.ne, m.cv4 = 1, object()  # OBB attribute surface (head.py:1011-1014)
    m.angle = (da.sigmoid() - 0.25) * math.pi  # OBB.forward stores the activated angle (head.py:1031-1034)
    _bind(m, "decode_bboxes", patch._wrap_decode_bboxes(_boom, head.obb_decode_bboxes))
    with torch.inference_mode():
        y = m._inference(dl)
        with no_host_sync():
            y = m._inference(dl)
        full = torch.cat([y, m.angle], 1)  # head.py:1038
    want = obb_forward_oracle(levels, ang, cfg.strides, cfg.nc)
    tol = 1e-5 * want.abs() + 1e-5 * cfg.imgsz
    assert not bool(((full.cpu() - want).abs() > tol).any())
    # 16-bit head: output dtype follows the input
    c2 = small_cfg(batch=2)
    lv16 = [lv.to(cuda_device) for lv in make_head_batch(c2, seed=5, dtype=torch.bfloat16)[0]]
    with torch.inference_mode():
        y16 = _detect(c2, cuda_device)._inference(lv16)
    assert y16.dtype == torch.bfloat16 and torch.equal(y16, head.decode_head(lv16, c2.strides, c2.nc))


def test_wrapper_hands_unsupported_configurations_to_the_reference(cuda_device):
    """reg_max != 16, export mode, autograd: the saved reference function runs (no RuntimeError, no lost grad)."""
    from ultralytics_pro_b200 import head, patch

    cfg = small_cfg(batch=1)
    dl = [lv.to(cuda_device) for lv in make_head_batch(cfg, seed=1)[0]]
    called = []
    m = StubDetect(cfg.nc, cfg.strides, cuda_device)
    _bind(m, "_inference", patch._wrap_inference(lambda self, x: called.append("ref") or "ref", head.detect_inference))
    with torch.inference_mode():
        m.reg_max = 8
        assert m._inference(dl) == "ref"
        m.reg_max, m.export = 16, True
        assert m._inference(dl) == "ref"
        m.export = False
        assert isinstance(m._inference(dl), torch.Tensor)
    with torch.enable_grad():
        assert m._inference(dl) == "ref"  # eval-mode saliency / attack tooling keeps its gradients
    assert called == ["ref"] * 3


# ------------------------------------------------------------------------------------------------------------------
# the side channel: _inference -> non_max_suppression reaches the fused kernels
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kw", [{}, {"agnostic": True}, {"classes": [0, 3, 17]}, {"multi_label": True, "conf_thres": 0.05},
                                {"max_det": 7}, {"return_idxs": True}])
def test_lazy_inference_then_nms_takes_the_fused_path_bit_exact(cuda_device, dtype, kw):
    from ultralytics_pro_b200 import head, lazy, nms, patch

    cfg = small_cfg(batch=3)
    levels, _ = make_head_batch(cfg, seed=37, dtype=dtype)
    dl = [lv.to(cuda_device) for lv in levels]
    m = _detect(cfg, cuda_device)
    nms_fn = patch._wrap_nms(_boom, nms.non_max_suppression)
    args = dict(conf_thres=0.25, iou_thres=0.7)
    args.update(kw)
    with torch.inference_mode():
        dense = m._inference(dl)
        want = nms_fn((dense, dl), **args)  # the tuple Detect.forward returns (head.py:124): two-call path
        lazy.ENABLED = True
        before = dict(lazy.STATS)
        y = m._inference(dl)
        assert isinstance(y, lazy.LazyDecoded) and y.shape == dense.shape and y.dtype == dense.dtype and y.is_cuda
        got = nms_fn((y, dl), **args)
    assert lazy.STATS["fused"] == before["fused"] + 1 and lazy.STATS["materialized"] == before["materialized"]
    if kw.get("return_idxs"):
        (got, gi), (want, wi) = got, want
        assert all(torch.equal(a, b) for a, b in zip(gi, wi))
    assert len(got) == 3 and sum(int(g.shape[0]) for g in got) > 0
    for g, w in zip(got, want):
        assert g.dtype == torch.float32 and torch.equal(g, w)
    if dtype == torch.float32 and not kw:
        ref_rows, _ = nms_oracle(dense.cpu(), 0.25, 0.7, nc=cfg.nc)
        assert_rows_equal(got, None, ref_rows, None, "fused drop-in vs oracle on our dense tensor")


def test_lazy_tensor_materialises_for_every_other_consumer(cuda_device):
    from ultralytics_pro_b200 import lazy, nms
    from ultralytics_pro_b200.head import decode_head, detect_postprocess

    cfg = small_cfg(batch=2)
    dl = [lv.to(cuda_device) for lv in make_head_batch(cfg, seed=41)[0]]
    dense = decode_head(dl, cfg.strides, cfg.nc)
    lazy.ENABLED = True
    with torch.inference_mode():
        m = _detect(cfg, cuda_device)
        y = m._inference(dl)
        assert torch.equal(torch.cat([y, y[:, :1]], 1), torch.cat([dense, dense[:, :1]], 1))  # Segment/Pose-style cat
        assert y.head_record() is None
        # a consumer the fused path does not serve (rotated / labels / nc mismatch) materialises, then runs from dense
        y = m._inference(dl)
        lab = [torch.tensor([[1.0, 10, 10, 30, 30]]), torch.zeros((0, 5))]
        got = nms.non_max_suppression(y, 0.25, 0.7, labels=lab)
        want = nms.non_max_suppression(dense, 0.25, 0.7, labels=lab)
        assert y.head_record() is None and all(torch.equal(a, b) for a, b in zip(got, want))
        # end2end heads, heads with riders and xyxy heads get the dense tensor right away
        assert type(_detect(cfg, cuda_device, end2end=True)._inference(dl)) is torch.Tensor
        seg = _detect(cfg, cuda_device)
        seg.nm = 32
        assert type(seg._inference(dl)) is torch.Tensor
        yx = _detect(cfg, cuda_device, xyxy=True)._inference(dl)  # lazy, but NMS must not take the xywh fused path
        assert isinstance(yx, lazy.LazyDecoded)
        assert torch.equal(yx.clone(), decode_head(dl, cfg.strides, cfg.nc, xyxy=True))
        # v10 flow: _inference -> permute -> postprocess (head.py:147-148) on a dense tensor
        e2e = _detect(cfg, cuda_device, end2end=True)
        out = detect_postprocess(e2e._inference(dl).permute(0, 2, 1), 20, cfg.nc)
        assert out.shape == (2, 20, 6)


# ------------------------------------------------------------------------------------------------------------------
# DFL.forward + decode_bboxes: the decode of YOLOEDetect.forward_lrpc (head.py:1777-1813)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_forward_lrpc_decode_chain(cuda_device, dtype):
    """forward_lrpc concatenates the levels itself and calls ``self.decode_bboxes(self.dfl(box), self.anchors.unsqueeze(0)) *
    self.strides`` (head.py:1804) - with DFL.forward and decode_bboxes bound, that line runs on the kernels."""
    from ultralytics_pro_b200 import head, patch

    cfg = small_cfg(batch=2)
    levels, _ = make_head_batch(cfg, seed=43, dtype=dtype)
    b = 2
    box_cpu = torch.cat([lv[:, :64].reshape(b, 64, -1) for lv in levels], 2)  # head.py:1797
    m = _detect(cfg, cuda_device)
    m.dfl = _bind(types.SimpleNamespace(c1=16), "forward", patch._wrap_dfl(_boom, head.dfl_forward))
    anc, srow = anchor_table(cfg.level_hw, cfg.strides, dtype)
    m.anchors, m.strides = anc.to(cuda_device), srow.to(cuda_device)  # (2, A) transposed views, like head.py:1794
    box = box_cpu.to(cuda_device)
    with torch.inference_mode():
        dist = m.dfl.forward(box)
        dbox = m.decode_bboxes(dist, m.anchors.unsqueeze(0)) * m.strides
        with no_host_sync():
            dist = m.dfl.forward(box)
            dbox = m.decode_bboxes(dist, m.anchors.unsqueeze(0)) * m.strides
    assert dist.dtype == dtype and dist.shape == (b, 4, cfg.anchors) and dbox.shape == (b, 4, cfg.anchors)
    want_dist = dfl_expect(box_cpu.float())  # block.py:250-253 in fp32 on the same (rounded) inputs
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    rel = (dist.cpu().float() - want_dist).abs() / want_dist.abs().clamp_min(1e-3)
    assert float(rel.max()) <= tol, f"DFL rel err {float(rel.max())}"
    # dist2bbox (tal.py:367-376) is exact arithmetic per op: bit-identical to the torch ops on OUR distances, any dtype
    d = dist.cpu()
    lt, rb = d.chunk(2, 1)
    p1, p2 = anc.unsqueeze(0) - lt, anc.unsqueeze(0) + rb
    want_box = torch.cat(((p1 + p2) / 2, p2 - p1), 1) * srow
    assert torch.equal(dbox.cpu(), want_box)
    # xyxy heads (head.py:189)
    m.xyxy = True
    with torch.inference_mode():
        assert torch.equal(m.decode_bboxes(dist, m.anchors.unsqueeze(0)).cpu(), torch.cat((p1, p2), 1))
        assert torch.equal(m.decode_bboxes(dist, m.anchors.unsqueeze(0), xywh=False).cpu(), torch.cat((p1, p2), 1))


def test_obb_decode_bboxes(cuda_device):
    from ultralytics_pro_b200 import head, patch

    cfg = small_cfg(batch=2, nc=15)
    g = torch.Generator().manual_seed(3)
    dist = torch.rand(2, 4, cfg.anchors, generator=g) * 12
    angle = (torch.rand(2, 1, cfg.anchors, generator=g) - 0.25) * math.pi
    anc, _ = anchor_table(cfg.level_hw, cfg.strides)
    m = StubDetect(cfg.nc, cfg.strides, cuda_device)
    m.angle = angle.to(cuda_device)
    _bind(m, "decode_bboxes", patch._wrap_decode_bboxes(_boom, head.obb_decode_bboxes))
    with torch.inference_mode():
        got = m.decode_bboxes(dist.to(cuda_device), anc.to(cuda_device).unsqueeze(0)).cpu()
    lt, rb = dist.split(2, dim=1)  # tal.py:397-403
    co, si = torch.cos(angle), torch.sin(angle)
    xf, yf = ((rb - lt) / 2).split(1, dim=1)
    want = torch.cat([torch.cat([xf * co - yf * si, xf * si + yf * co], 1) + anc.unsqueeze(0), lt + rb], 1)
    assert torch.equal(got[:, 2:], want[:, 2:])  # w, h: exact
    assert float((got[:, :2] - want[:, :2]).abs().max()) <= 1e-5 * 16  # cos / sin of two libms


# ------------------------------------------------------------------------------------------------------------------
# Pose.kpts_decode
# ------------------------------------------------------------------------------------------------------------------
def test_bound_kpts_decode_both_grid_sources(cuda_device):
    from ultralytics_pro_b200 import head, patch

    cfg = small_cfg("pose", nc=1, batch=2)
    levels, _ = make_head_batch(cfg, seed=47)
    dl = [lv.to(cuda_device) for lv in levels]
    g = torch.Generator().manual_seed(47)
    kp = torch.randn(2, 51, cfg.anchors, generator=g)
    want = ro.kpts_decode_oracle(kp, cfg.level_hw, cfg.strides, (17, 3))
    m = _detect(cfg, cuda_device)
    m.kpt_shape, m.nk = (17, 3), 51
    _bind(m, "kpts_decode", patch._wrap_kpts(_boom, head.pose_kpts_decode))
    kd = kp.to(cuda_device)
    with torch.inference_mode():
        m._inference(dl)  # Pose.forward runs Detect.forward first (head.py:1249)
        got = m.kpts_decode(2, kd)
        with no_host_sync():
            got = m.kpts_decode(2, kd.clone())
    g4, w4 = got.cpu().view(2, 17, 3, -1), want.view(2, 17, 3, -1)
    assert torch.equal(g4[:, :, :2], w4[:, :, :2]) and float((g4[:, :, 2] - w4[:, :, 2]).abs().max()) < 1e-6
    # a module whose _inference ran through the reference: grids recovered from the cached anchor rows, once
    m2 = StubDetect(cfg.nc, cfg.strides, cuda_device)
    m2.kpt_shape, m2.nk = (17, 3), 51
    anc, srow = anchor_table(cfg.level_hw, cfg.strides)
    m2.anchors, m2.strides, m2.shape = anc.to(cuda_device), srow.to(cuda_device), dl[0].shape
    _bind(m2, "kpts_decode", patch._wrap_kpts(_boom, head.pose_kpts_decode))
    got2 = m2.kpts_decode(2, kd)
    with no_host_sync():
        got3 = m2.kpts_decode(2, kd.clone())
    assert torch.equal(got2, got) and torch.equal(got3, got)


# ------------------------------------------------------------------------------------------------------------------
# validator, TorchNMS, ops and mask wrappers: the CUDA branch of each
# ------------------------------------------------------------------------------------------------------------------
def test_bound_validator_methods(cuda_device):
    from ultralytics_pro_b200 import patch, val

    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(60, 2, generator=g) * 400
    boxes = torch.cat([xy, xy + torch.rand(60, 2, generator=g) * 200 + 4], 1)
    gt = boxes[:20] + torch.randn(20, 4, generator=g) * 3
    pcls, gcls = torch.randint(0, 3, (60,), generator=g).float(), torch.randint(0, 3, (20,), generator=g).float()
    iouv = torch.linspace(0.5, 0.95, 10)
    iou = ro.box_iou_oracle(gt.numpy(), boxes.numpy())
    want = ro.match_predictions_oracle(pcls.numpy(), gcls.numpy(), iou, iouv.tolist())
    me = types.SimpleNamespace(iouv=iouv.to(dev), niou=10)  # validator.py:148 keeps iouv on the device
    _bind(me, "match_predictions", patch._wrap_match(_boom, val.match_predictions))
    _bind(me, "_process_batch", patch._wrap_process_batch(_boom, val.process_batch))
    iou_d, pcls_d, gcls_d = torch.from_numpy(iou).to(dev), pcls.to(dev), gcls.to(dev)
    got = me.match_predictions(pcls_d, gcls_d, iou_d)
    with no_host_sync():  # the IoU levels were read back once, on the first call
        got = me.match_predictions(pcls_d, gcls_d, iou_d)
    assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), want)
    tp = me._process_batch({"bboxes": boxes.to(dev), "cls": pcls.to(dev)}, {"bboxes": gt.to(dev), "cls": gcls.to(dev)})["tp"]
    assert np.array_equal(tp, want)
    # use_scipy / CPU inputs go to the reference
    ref_called = []
    _bind(me, "match_predictions", patch._wrap_match(lambda self, *a: ref_called.append(1), val.match_predictions))
    me.match_predictions(pcls.to(dev), gcls.to(dev), iou_d, True)
    me.match_predictions(pcls, gcls, torch.from_numpy(iou))
    assert len(ref_called) == 2


def test_bound_torchnms_ops_and_masks(cuda_device):
    from ultralytics_pro_b200 import nms, ops, patch
    from oracle.postproc_oracle import box_iou_matrix, fast_nms, greedy_nms, probiou_matrix

    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    xy = torch.rand(300, 2, generator=g) * 300
    boxes = torch.cat([xy, xy + torch.rand(300, 2, generator=g) * 80 + 2], 1)
    scores = torch.rand(300, generator=g)
    f_nms = patch._wrap_static(_boom, nms.TorchNMS.nms).__func__
    f_fast = patch._wrap_static(_boom, nms.TorchNMS.fast_nms).__func__
    f_bat = patch._wrap_static(_boom, nms.TorchNMS.batched_nms).__func__
    assert torch.equal(f_nms(boxes.to(dev), scores.to(dev), 0.5).cpu(), greedy_nms(boxes, scores, 0.5, impl="torchnms"))
    assert torch.equal(f_fast(boxes.to(dev), scores.to(dev), 0.5, iou_func=nms.box_iou).cpu(), fast_nms(boxes, scores, 0.5, iou="box"))
    idxs = torch.randint(0, 4, (300,), generator=g)
    off = idxs.to(boxes) * (boxes.max() + 1)
    assert torch.equal(f_bat(boxes.to(dev), scores.to(dev), idxs.to(dev), 0.5).cpu(), greedy_nms(boxes + off[:, None], scores, 0.5, impl="torchnms"))
    # metrics.box_iou / batch_probiou as free functions (they used to be markers that raised)
    got = nms.box_iou(boxes[:50].to(dev), boxes[50:120].to(dev)).cpu()
    assert torch.equal(got, box_iou_matrix(boxes[:50], boxes[50:120]))
    obb = torch.cat([xy, torch.rand(300, 2, generator=g) * 80 + 2, (torch.rand(300, 1, generator=g) - 0.25) * math.pi], 1)
    gotp = nms.batch_probiou(obb[:40].to(dev), obb[40:100].to(dev)).cpu()
    assert float((gotp - probiou_matrix(obb[:40], obb[40:100])).abs().max()) < 2e-6
    # ops wrappers
    rows = torch.cat([boxes[:40], scores[:40, None], torch.zeros(40, 1)], 1)
    f_scale = patch._wrap_ops(_named(_boom, "scale_boxes"), ops.scale_boxes, 1)
    f_clip = patch._wrap_ops(_named(_boom, "clip_boxes"), ops.clip_boxes, 0)
    r = rows.to(dev)
    with no_host_sync():
        f_scale((640, 640), r[:, :4], (480, 600, 3))
    assert np.array_equal(r[:, :4].cpu().numpy(), ro.scale_boxes_oracle((640, 640), rows[:, :4].numpy(), (480, 600, 3)))
    r = rows.to(dev)
    f_clip(r[:, :4], (100, 120))
    assert np.array_equal(r[:, :4].cpu().numpy(), ro.clip_boxes_oracle(rows[:, :4].numpy().copy(), (100, 120)))
    kp = torch.rand(40, 17, 3, generator=g) * 600
    f_coords = patch._wrap_ops(_named(_boom, "scale_coords"), ops.scale_coords, 1)
    kd = kp.to(dev)
    f_coords((640, 640), kd, (480, 600))
    assert np.array_equal(kd.cpu().numpy(), ro.scale_coords_oracle((640, 640), kp.numpy().copy(), (480, 600)))
    # masks
    protos, coef = torch.randn(32, 40, 40, generator=g), torch.randn(60, 32, generator=g)
    mb = boxes[:60].clamp(0, 160)
    f_mask = patch._wrap_masks(_named(_boom, "process_mask"), ops.process_mask)
    m_ref, m_val = ro.process_mask_oracle(protos, coef, mb, (160, 160), upsample=True)
    got = f_mask(protos.to(dev), coef.to(dev), mb.to(dev), (160, 160), upsample=True).cpu()
    bad = got != m_ref
    assert not bool(bad.any()) or float(m_val[bad].abs().max()) < 1e-4


def _named(fn, name):
    def f(*a, **k):
        return fn(*a, **k)

    f.__name__ = name
    return f
