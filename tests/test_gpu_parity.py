"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): decode within 1e-5 relative for fp32 (implemented as |d| <= 1e-5*|ref| + 1e-5*imgsz,
SURVEY.md Appendix B-7) / 1e-2 for bf16; NMS kept indices, counts and row values bit-exact when both sides are fed
the same decoded tensor.
"""
import pytest
import torch

from oracle.postproc_oracle import decode_oracle, nms_oracle, obb_forward_oracle
from tests.helpers import assert_rows_equal, dense_from_oracle, make_scores_unique, small_cfg
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

pytestmark = pytest.mark.gpu


def _to(dev, levels, ang=None):
    return [lv.to(dev) for lv in levels], (ang.to(dev) if ang is not None else None)


# ------------------------------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,batch", [("c1_v8n_640_b1", 1), ("c2_v8x_640_b64", 3), ("c4_p6_1280_b16", 2)])
def test_decode_dense_fp32(cuda_device, name, batch):
    from ultralytics_pro_b200.head import decode_head

    cfg = CONFIGS[name]
    levels, _ = make_head_batch(cfg, batch=batch, seed=11)
    want = decode_oracle(levels, cfg.strides, cfg.nc)
    got = decode_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc).cpu()
    assert got.shape == want.shape and got.dtype == want.dtype
    tol = 1e-5 * want.abs() + 1e-5 * cfg.imgsz
    bad = (got - want).abs() > tol
    assert not bool(bad[:, :4].any()), f"box mismatch max {float((got - want)[:, :4].abs().max())}"
    # scores and w/h are well conditioned: pure relative bound
    rel = ((got - want).abs() / want.abs().clamp_min(1e-30))
    assert float(rel[:, 4:].max()) < 1e-5, f"score rel err {float(rel[:, 4:].max())}"
    assert float(rel[:, 2:4].max()) < 1e-5, f"wh rel err {float(rel[:, 2:4].max())}"


def test_decode_dense_xyxy(cuda_device):
    from ultralytics_pro_b200.head import decode_head

    cfg = small_cfg(batch=2)
    levels, _ = make_head_batch(cfg, seed=5)
    want = decode_oracle(levels, cfg.strides, cfg.nc, xyxy=True)
    got = decode_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc, xyxy=True).cpu()
    assert float((got - want).abs().max()) <= 1e-5 * cfg.imgsz


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_decode_dense_half(cuda_device, dtype):
    from ultralytics_pro_b200.head import decode_head

    cfg = CONFIGS["c2_v8x_640_b64"]
    levels, _ = make_head_batch(cfg, batch=2, seed=13, dtype=dtype)
    got = decode_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc).cpu()
    assert got.dtype == dtype
    exact = decode_oracle([lv.float() for lv in levels], cfg.strides, cfg.nc)  # fp32 math on the same (rounded) inputs
    ref = decode_oracle(levels, cfg.strides, cfg.nc).float()                   # the reference's own low-precision chain
    tol = 1e-2 * exact.abs() + 1e-2
    err_ours = (got.float() - exact).abs()
    assert not bool((err_ours > tol).any()), f"max err {float(err_ours.max())}"
    # never worse than the reference's own per-op-rounded chain, in aggregate
    assert float(err_ours.mean()) <= float((ref - exact).abs().mean()) * 1.05 + 1e-6


def test_decode_obb(cuda_device):
    from ultralytics_pro_b200.head import decode_head

    cfg = CONFIGS["c5_obb_1024_b16"]
    levels, ang = make_head_batch(cfg, batch=2, seed=17)
    want = obb_forward_oracle(levels, ang, cfg.strides, cfg.nc)
    dl, da = _to(cuda_device, levels, ang)
    got = decode_head(dl, cfg.strides, cfg.nc, angle=da, angle_is_logit=True, append_angle=True).cpu()
    assert got.shape == want.shape
    tol = 1e-5 * want.abs() + 1e-5 * cfg.imgsz
    assert not bool(((got - want).abs() > tol).any()), f"max {float((got - want).abs().max())}"


def test_decode_scalar_path_odd_grid(cuda_device):
    """Level sizes that are not multiples of the vector width take the VEC=1 kernel."""
    from ultralytics_pro_b200.head import decode_head

    torch.manual_seed(3)
    nc = 7
    levels = [torch.randn(2, 64 + nc, 9, 7), torch.randn(2, 64 + nc, 5, 3)]
    want = decode_oracle(levels, (8, 16), nc)
    got = decode_head(_to(cuda_device, levels)[0], (8, 16), nc).cpu()
    assert float((got - want).abs().max()) <= 1e-3


# ------------------------------------------------------------------------------------------------------------------
# non_max_suppression on the same decoded tensor: bit-exact
# ------------------------------------------------------------------------------------------------------------------
CASES = [
    # name, batch, overrides
    ("c2_v8x_640_b64", 4, {}),
    ("c2_v8x_640_b64", 2, {"agnostic": True}),
    ("c2_v8x_640_b64", 2, {"classes": [0, 3, 17, 42, 79]}),
    ("c2_v8x_640_b64", 2, {"max_det": 20}),
    ("c2_v8x_640_b64", 2, {"conf": 0.05, "iou": 0.45}),
    ("c4_p6_1280_b16", 2, {}),
    ("c4_p6_1280_b16", 2, {"agnostic": True}),
    ("c3_val_stress_b32", 2, {}),
    ("c3_val_stress_b32", 1, {"max_nms": 5000}),
    # the benchmarked sizes (bench.py: C2 at B=64; config sweep: C4 at B=16, C3 at B=8 of 32): the batch-dependent machinery -
    # GPU-wide octet list, one suppression CTA per image on 64 SMs at once, the prefix select of the 29 k-row images
    ("c2_v8x_640_b64", 64, {}),
    ("c4_p6_1280_b16", 16, {}),
    ("c3_val_stress_b32", 8, {}),
    # the prefix (radix select of the best rows) is not enough: max_det larger than the prefix can hold -> full sort + walk
    ("c3_val_stress_b32", 1, {"max_det": 3000}),
    # ... and a prefix that is tried but does not yield max_det kept rows (aggressive threshold): prefix, then full walk
    ("c3_val_stress_b32", 1, {"iou": 0.1, "max_det": 1000}),
]


@pytest.mark.parametrize("name,batch,over", CASES)
def test_nms_from_dense_bitexact(cuda_device, name, batch, over):
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = CONFIGS[name]
    _, _, y = dense_from_oracle(cfg, batch, seed=21)
    kw = dict(conf=cfg.conf, iou=cfg.iou, multi_label=cfg.multi_label, agnostic=cfg.agnostic, max_det=cfg.max_det,
              max_nms=cfg.max_nms, classes=None)
    kw.update(over)
    if kw["max_nms"] < 30000:
        y = make_scores_unique(y, cfg.nc, kw["conf"])
    want, want_idx = nms_oracle(y, kw["conf"], kw["iou"], classes=kw["classes"], agnostic=kw["agnostic"],
                                multi_label=kw["multi_label"], max_det=kw["max_det"], nc=cfg.nc, max_nms=kw["max_nms"])
    got, got_idx = non_max_suppression(y.to(cuda_device), kw["conf"], kw["iou"], classes=kw["classes"],
                                       agnostic=kw["agnostic"], multi_label=kw["multi_label"], max_det=kw["max_det"],
                                       nc=cfg.nc, max_nms=kw["max_nms"], return_idxs=True)
    assert sum(w.shape[0] for w in want) > 0
    assert_rows_equal(got, got_idx, want, want_idx, f"{name} {over}")


def test_nms_rotated_matches_oracle(cuda_device):
    """ProbIoU goes through cos/sin/log/exp whose last ulp differs between libm (CPU) and CUDA: kept sets must match
    except for pairs within 1e-6 of the threshold; here we demand equality and report the margin."""
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = CONFIGS["c5_obb_1024_b16"]
    for batch, conf, iou in ((3, cfg.conf, cfg.iou), (16, cfg.conf, cfg.iou), (2, 0.01, 0.3)):  # 16 = the benchmarked batch
        _, _, y = dense_from_oracle(cfg, batch, seed=23)
        y = make_scores_unique(y, cfg.nc, conf)
        want, want_idx = nms_oracle(y, conf, iou, nc=cfg.nc, rotated=True)
        got, got_idx = non_max_suppression(y.to(cuda_device), conf, iou, nc=cfg.nc, rotated=True, return_idxs=True)
        assert sum(w.shape[0] for w in want) > 0
        assert_rows_equal(got, got_idx, want, want_idx, f"obb B={batch} conf={conf} iou={iou}")


def test_fast_nms_cheap_rejection_is_safe_on_adversarial_boxes(cuda_device):
    """The cluster Fast-NMS kernel skips the full ProbIoU formula for pairs its cheap bound proves to be below the threshold
    (|d|^2 > K (T1 + T2), ypb_nms.cu).  Boxes chosen to stress that bound: extreme aspect ratios (irregular -> never
    skipped), needle boxes at 45 degrees, tiny and huge boxes, heavy overlaps and near-touching neighbours, several
    thresholds.  Kept indices must equal the oracle's (the N x N ProbIoU matrix of metrics.py:251-284)."""
    from oracle.postproc_oracle import fast_nms
    from ultralytics_pro_b200.nms import TorchNMS, batch_probiou

    g = torch.Generator().manual_seed(77)
    n = 1500
    centres = torch.rand(40, 2, generator=g) * 900
    own = torch.randint(0, 40, (n,), generator=g)
    xy = centres[own] + torch.randn(n, 2, generator=g) * torch.rand(n, 1, generator=g) * 40
    w = torch.exp(torch.rand(n, generator=g) * 9 - 2)            # 0.13 .. 1100 px
    ar = torch.exp((torch.rand(n, generator=g) - 0.5) * 12)      # aspect ratios 1/400 .. 400
    h = (w / ar).clamp(1e-3, 5e3)
    ang = (torch.rand(n, generator=g) - 0.25) * 3.14159265
    ang[::7] = 0.78539816                                        # needles on the diagonal: C ~ (a - b) / 2
    obb = torch.cat([xy, w[:, None], h[:, None], ang[:, None]], 1)
    obb[5::50, :2] = obb[4::50, :2]                              # exact duplicates of the centre
    scores = torch.rand(n, generator=g)
    for thr in (0.7, 0.45, 0.1, 0.01, 0.9):
        want = fast_nms(obb, scores, thr, "probiou")
        got = TorchNMS.fast_nms(obb.to(cuda_device), scores.to(cuda_device), thr, iou_func=batch_probiou).cpu()
        if not torch.equal(got, want):
            # a differing row must be a borderline pair (CUDA and SLEEF transcendentals differ in the last ulp), never a skipped one
            from oracle.postproc_oracle import probiou_matrix, stable_desc_order

            order = stable_desc_order(scores)
            m = probiou_matrix(obb[order], obb[order]).triu_(1)
            diff = set(got.tolist()) ^ set(want.tolist())
            pos = {int(v): i for i, v in enumerate(order.tolist())}
            for idx in diff:
                col = m[:, pos[idx]]
                assert float((col - thr).abs().min()) < 1e-5, f"thr={thr}: row {idx} differs without a borderline pair"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_nms_from_dense_half_inputs(cuda_device, dtype):
    """Low-precision predictions: threshold cast, in-dtype xywh->xyxy and fp32 promotion of the rows (nms.py:76,86,116)."""
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = CONFIGS["c2_v8x_640_b64"]
    _, _, y = dense_from_oracle(cfg, 2, seed=29)
    y = y.to(dtype)
    want, want_idx = nms_oracle(y, cfg.conf, cfg.iou, nc=cfg.nc)
    got, got_idx = non_max_suppression(y.to(cuda_device), cfg.conf, cfg.iou, nc=cfg.nc, return_idxs=True)
    assert got[0].dtype == torch.float32
    assert_rows_equal(got, got_idx, want, want_idx, str(dtype))


def test_nms_extras_and_noncontiguous(cuda_device):
    """Mask-coefficient style extras ride through (nms.py:74,112) and permuted inputs work (models/nas/predict.py:54)."""
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = small_cfg(imgsz=320, batch=3)
    _, _, y = dense_from_oracle(cfg, 3, seed=31)
    torch.manual_seed(1)
    y = torch.cat((y, torch.randn(3, 32, y.shape[2])), 1)
    want, want_idx = nms_oracle(y, 0.25, 0.7, nc=cfg.nc)
    yd = y.to(cuda_device).permute(0, 2, 1).contiguous().permute(0, 2, 1)  # (B, C, A) view with stride_a = C
    assert not yd.is_contiguous()
    got, got_idx = non_max_suppression(yd, 0.25, 0.7, nc=cfg.nc, return_idxs=True)
    assert got[0].shape[1] == 6 + 32
    assert_rows_equal(got, got_idx, want, want_idx, "extras")


def test_nms_empty_and_edge(cuda_device):
    from ultralytics_pro_b200.nms import non_max_suppression

    y = torch.zeros(2, 84, 64, device=cuda_device)
    out, idx = non_max_suppression(y, 0.25, 0.7, return_idxs=True)
    assert [o.shape for o in out] == [(0, 6), (0, 6)] and all(i.numel() == 0 for i in idx)
    # one image empty, one with a single box; tuple input accepted (nms.py:61)
    y[1, :4, 5] = torch.tensor([50.0, 60.0, 20.0, 10.0], device=cuda_device)
    y[1, 4 + 3, 5] = 0.9
    out = non_max_suppression((y, None), 0.25, 0.7)
    assert out[0].shape == (0, 6) and out[1].shape == (1, 6)
    assert out[1][0].tolist() == [40.0, 55.0, 60.0, 65.0, pytest.approx(0.9), 3.0]
    with pytest.raises(AssertionError):
        non_max_suppression(y, 1.5, 0.7)
    with pytest.raises(RuntimeError):
        non_max_suppression(y.cpu(), 0.25, 0.7)


def test_nms_iou_equal_threshold_and_degenerate(cuda_device):
    """IoU == thr keeps both (strict >, nms.py:292-294); zero-area duplicates are both kept (0/0 = NaN); thr=0.6 hits the
    float-vs-double threshold corner (torchvision CPU suppresses IoU == float32(0.6))."""
    from ultralytics_pro_b200.nms import non_max_suppression

    def run(boxes_xyxy, scores, thr):
        n = len(scores)
        y = torch.zeros(1, 5, n)
        b = torch.tensor(boxes_xyxy, dtype=torch.float32)
        y[0, 0], y[0, 1] = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
        y[0, 2], y[0, 3] = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        y[0, 4] = torch.tensor(scores)
        want, _ = nms_oracle(y, 0.1, thr, nc=1)
        got = non_max_suppression(y.to(cuda_device), 0.1, thr, nc=1)
        assert_rows_equal(got, None, want, None, f"thr={thr}")
        return got[0].shape[0]

    assert run([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.5) == 2
    assert run([[1, 1, 1, 1], [1, 1, 1, 1]], [0.9, 0.8], 0.5) == 2
    assert run([[0, 0, 10, 10], [0, 0, 10, 6]], [0.9, 0.8], 0.6) == 1


# ------------------------------------------------------------------------------------------------------------------
# fused path == dense decode followed by NMS, bit for bit
# ------------------------------------------------------------------------------------------------------------------
FUSED = [
    ("c2_v8x_640_b64", 4, torch.float32),
    ("c2_v8x_640_b64", 2, torch.bfloat16),
    ("c2_v8x_640_b64", 2, torch.float16),
    ("c3_val_stress_b32", 2, torch.float32),
    ("c4_p6_1280_b16", 2, torch.float32),
    ("c5_obb_1024_b16", 2, torch.float32),
    # the benchmarked sizes
    ("c2_v8x_640_b64", 64, torch.float32),
    ("c2_v8x_640_b64", 64, torch.bfloat16),
    ("c3_val_stress_b32", 8, torch.float32),
    ("c4_p6_1280_b16", 16, torch.float32),
    ("c5_obb_1024_b16", 16, torch.float32),
]


@pytest.mark.parametrize("name,batch,dtype", FUSED)
def test_fused_equals_two_call(cuda_device, name, batch, dtype):
    from ultralytics_pro_b200.head import decode_head, postprocess_from_head
    from ultralytics_pro_b200.nms import non_max_suppression

    cfg = CONFIGS[name]
    levels, ang = make_head_batch(cfg, batch=batch, seed=37, dtype=dtype)
    dl, da = _to(cuda_device, levels, ang)
    if cfg.rotated:
        y = decode_head(dl, cfg.strides, cfg.nc, angle=da, angle_is_logit=True, append_angle=True)
    else:
        y = decode_head(dl, cfg.strides, cfg.nc)
    two, two_idx = non_max_suppression(y, cfg.conf, cfg.iou, nc=cfg.nc, multi_label=cfg.multi_label,
                                       rotated=cfg.rotated, return_idxs=True)
    one, one_idx = postprocess_from_head(dl, cfg.strides, cfg.nc, cfg.conf, cfg.iou, multi_label=cfg.multi_label,
                                         angle_logits=da, return_idxs=True)
    assert sum(t.shape[0] for t in two) > 0
    assert_rows_equal(one, one_idx, [t.cpu() for t in two], [t.cpu() for t in two_idx], f"fused {name} {dtype}")


@pytest.mark.parametrize("name,batch,dtype", [("c2_v8x_640_b64", 64, torch.float32), ("c2_v8x_640_b64", 5, torch.bfloat16),
                                              ("c2_v8x_640_b64", 3, torch.float16), ("c4_p6_1280_b16", 3, torch.float32),
                                              ("c5_obb_1024_b16", 4, torch.float32), ("c3_val_stress_b32", 2, torch.float32)])
def test_tma_scan_kernel_equals_ldg_scan_kernel(cuda_device, name, batch, dtype):
    """The two forms of the class-scan kernel (ypb_nms_params.scan_kernel): one-wave register-staged LDG grid vs the
    persistent TMA-fed ring (cp.async.bulk.tensor + mbarrier, ypb_scan_tma.cu) must hand identical candidates to the rest of
    the fused path - rows, kept anchors and counts bit for bit, partial last tiles of a level and 15-class heads included."""
    from ultralytics_pro_b200.pipeline import HeadPostProcessor

    cfg = CONFIGS[name]
    levels, ang = make_head_batch(cfg, batch=batch, seed=51, dtype=dtype)
    dl, da = _to(cuda_device, levels, ang)
    res = {}
    for kern in ("ldg", "tma"):
        pp = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, multi_label=cfg.multi_label, rotated=cfg.rotated,
                               max_det=cfg.max_det, max_nms=cfg.max_nms, scan_kernel=kern)
        rows, idx = pp(dl, da, return_idxs=True)
        res[kern] = (rows, idx, pp.last.cand.clone())
    assert sum(r.shape[0] for r in res["ldg"][0]) > 0
    assert torch.equal(res["ldg"][2], res["tma"][2]), "candidate counts differ"
    for a, b, ia, ib in zip(res["ldg"][0], res["tma"][0], res["ldg"][1], res["tma"][1]):
        assert torch.equal(a, b) and torch.equal(ia, ib)


def test_tma_scan_falls_back_on_geometries_outside_its_envelope(cuda_device):
    """Odd grids (not 16-byte vectorisable) silently take the LDG kernel: same results as the reference path."""
    from ultralytics_pro_b200.head import decode_head
    from ultralytics_pro_b200.nms import non_max_suppression
    from ultralytics_pro_b200.pipeline import HeadPostProcessor

    torch.manual_seed(5)
    nc = 7
    levels = [torch.randn(2, 64 + nc, 9, 7) * 2, torch.randn(2, 64 + nc, 5, 3) * 2]
    dl = [lv.to(cuda_device) for lv in levels]
    pp = HeadPostProcessor(nc, (8, 16), 0.25, 0.7, scan_kernel="tma")
    got = pp(dl)
    want = non_max_suppression(decode_head(dl, (8, 16), nc), 0.25, 0.7)
    assert sum(w.shape[0] for w in want) > 0 and all(torch.equal(g, w) for g, w in zip(got, want))


def _decision_margins(y_img: torch.Tensor, nc: int, conf: float, iou_thr: float, max_wh: float = 7680.0):
    """How far the oracle's decisions on one decoded image are from flipping: (min |score - conf| over the anchors' best
    scores, min |IoU - thr| over the comparisons the greedy walk actually makes - a kept row against every row still alive
    below it, nms.py:276-294)."""
    import numpy as np

    best, cls = y_img[4:4 + nc].max(0)
    conf_margin = float((best - conf).abs().min())
    cand = torch.nonzero(best > conf).squeeze(1)
    order = cand[torch.sort(best[cand], descending=True, stable=True).indices]
    b = y_img[:4, order].t().numpy().astype(np.float32)
    off = (cls[order].float() * max_wh).numpy().astype(np.float32)[:, None]
    xyxy = np.concatenate([b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2], 1) + off
    area = (xyxy[:, 2] - xyxy[:, 0]) * (xyxy[:, 3] - xyxy[:, 1])
    dead = np.zeros(len(order), bool)
    iou_margin = 1.0
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(len(order)):
            if dead[i] or i + 1 == len(order):
                continue
            w = np.maximum(0, np.minimum(xyxy[i, 2], xyxy[i + 1:, 2]) - np.maximum(xyxy[i, 0], xyxy[i + 1:, 0]))
            h = np.maximum(0, np.minimum(xyxy[i, 3], xyxy[i + 1:, 3]) - np.maximum(xyxy[i, 1], xyxy[i + 1:, 1]))
            inter = w * h
            iou = inter / (area[i] + area[i + 1:] - inter)
            live = ~dead[i + 1:] & (inter > 0)
            if live.any():
                iou_margin = min(iou_margin, float(np.abs(iou[live] - iou_thr).min()))
            dead[i + 1:] |= iou > iou_thr
    return conf_margin, iou_margin


def test_fused_against_oracle_end_to_end(cuda_device):
    """Our decode + NMS against the oracle's decode + NMS.  The two decodes agree to ~1e-4 px, which
    moves a score by < 1e-6 and an IoU by < 1e-5; an image is compared only if none of the oracle's decisions sits inside
    those margins (score within 1e-5 of conf; an IoU the greedy walk evaluates within 1e-5 of the threshold - nms.py:76,
    292-294).  On every remaining image the kept anchors must be IDENTICAL and the row values within the decode tolerance."""
    from ultralytics_pro_b200.head import postprocess_from_head

    cfg = CONFIGS["c2_v8x_640_b64"]
    nb = 16
    levels, _ = make_head_batch(cfg, batch=nb, seed=41)
    y = decode_oracle(levels, cfg.strides, cfg.nc)
    want, want_idx = nms_oracle(y, cfg.conf, cfg.iou, nc=cfg.nc)
    got, got_idx = postprocess_from_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc, cfg.conf, cfg.iou, return_idxs=True)
    compared = 0
    for b in range(nb):
        conf_margin, iou_margin = _decision_margins(y[b], cfg.nc, cfg.conf, cfg.iou)
        if conf_margin < 1e-5 or iou_margin < 1e-5:
            continue
        compared += 1
        gi, wi = got_idx[b].cpu(), want_idx[b]
        assert gi.shape == wi.shape and torch.equal(gi, wi), f"image {b}: kept anchors differ (margins {conf_margin:.2e}, {iou_margin:.2e})"
        assert float((got[b].cpu() - want[b]).abs().max()) <= 1e-5 * cfg.imgsz + 1e-5 * float(want[b].abs().max())
    assert compared >= nb // 2, f"only {compared} of {nb} images had margin-free decisions"


# ------------------------------------------------------------------------------------------------------------------
# TorchNMS mirror, self tests
# ------------------------------------------------------------------------------------------------------------------
def test_torchnms_mirror(cuda_device):
    from oracle.postproc_oracle import fast_nms, greedy_nms
    from ultralytics_pro_b200.nms import TorchNMS, batch_probiou, box_iou

    torch.manual_seed(7)
    n = 700
    xy = torch.rand(n, 2) * 300
    wh = torch.rand(n, 2) * 80 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = torch.rand(n)
    want = greedy_nms(boxes, scores, 0.5, "plain")
    got = TorchNMS.nms(boxes.to(cuda_device), scores.to(cuda_device), 0.5).cpu()
    assert torch.equal(got, want)
    want = fast_nms(boxes, scores, 0.5, "boxiou")
    got = TorchNMS.fast_nms(boxes.to(cuda_device), scores.to(cuda_device), 0.5, iou_func=box_iou).cpu()
    assert torch.equal(got, want)
    obb = torch.cat((xy, wh, (torch.rand(n, 1) - 0.25) * 3.14159), 1)
    want = fast_nms(obb, scores, 0.3, "probiou")
    got = TorchNMS.fast_nms(obb.to(cuda_device), scores.to(cuda_device), 0.3, iou_func=batch_probiou).cpu()
    assert torch.equal(got, want)
    idxs = torch.randint(0, 5, (n,))
    got = TorchNMS.batched_nms(boxes.to(cuda_device), scores.to(cuda_device), idxs.to(cuda_device), 0.5).cpu()
    off = idxs.to(boxes) * (boxes.max() + 1)
    want = greedy_nms(boxes + off[:, None], scores, 0.5, "plain")
    assert torch.equal(got, want)
    # large n: radix-sort + multi-chunk path
    n = 9000
    xy = torch.rand(n, 2) * 2000
    wh = torch.rand(n, 2) * 60 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = torch.rand(n)
    want = greedy_nms(boxes, scores, 0.5, "torchvision")
    got = TorchNMS.nms(boxes.to(cuda_device), scores.to(cuda_device), 0.5).cpu()
    assert torch.equal(got, want)
    assert TorchNMS.nms(torch.zeros(0, 4, device=cuda_device), torch.zeros(0, device=cuda_device), 0.5).numel() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_sigmoid_monotone_selftest(cuda_device, dtype):
    """The single-label fast filter relies on round_T(sigmoid(x)) being monotone: checked over all 2^32 inputs."""
    from ultralytics_pro_b200 import _cabi

    lib = _cabi.load()
    v = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    _cabi.check(lib.ypb_selftest_sigmoid_monotone(_cabi.dtype_code(dtype), v.data_ptr(), _cabi.stream_ptr(cuda_device)), "selftest")
    assert int(v.item()) == 0


# ------------------------------------------------------------------------------------------------------------------
# fallback and corner paths of the suppression kernel
# ------------------------------------------------------------------------------------------------------------------
def _random_decoded(n_anchors, nc, batch, seed, span=600.0, extra=0, crowd_cls=None):
    g = torch.Generator().manual_seed(seed)
    y = torch.zeros(batch, 4 + nc + extra, n_anchors)
    cx = torch.rand(batch, n_anchors, generator=g) * span
    cy = torch.rand(batch, n_anchors, generator=g) * span
    w = torch.rand(batch, n_anchors, generator=g) * 80 + 4
    h = torch.rand(batch, n_anchors, generator=g) * 80 + 4
    y[:, 0], y[:, 1], y[:, 2], y[:, 3] = cx, cy, w, h
    sc = torch.rand(batch, nc, n_anchors, generator=g) * 0.2
    hot = torch.randint(0, nc, (batch, n_anchors), generator=g) if crowd_cls is None else torch.full((batch, n_anchors), crowd_cls)
    sc.scatter_(1, hot.unsqueeze(1), torch.rand(batch, 1, n_anchors, generator=g))
    y[:, 4:4 + nc] = sc
    if extra:
        y[:, 4 + nc:] = torch.randn(batch, extra, n_anchors, generator=g)
    return y


@pytest.mark.parametrize("case", ["single_class", "crowded_class", "wide_span", "big_max_det", "many_classes", "tiny_nc2"])
def test_suppression_fallback_paths(cuda_device, case):
    """Every precondition of the class-wise walk, violated in turn, must land on the dense walk with identical results."""
    from ultralytics_pro_b200.nms import non_max_suppression

    kw = dict(conf=0.3, iou=0.5, max_det=300)
    if case == "single_class":      # nc = 1: one class holds every row (> 256) -> dense walk
        y = _random_decoded(3000, 1, 2, 1)
    elif case == "crowded_class":   # 80 classes but one of them holds ~1500 rows
        y = _random_decoded(3000, 80, 2, 2, crowd_cls=7)
    elif case == "wide_span":       # coordinates span more than max_wh: classes may interact through the offset
        y = _random_decoded(1500, 4, 2, 3, span=20000.0)
        y[:, 2:4] *= 60             # boxes thousands of pixels wide, so cross-class overlaps really occur
    elif case == "big_max_det":     # kept list larger than the shared-memory list
        y = _random_decoded(6000, 1, 1, 4, span=4000.0)
        kw.update(max_det=2500, iou=0.3)
    elif case == "many_classes":    # more classes than histogram bins -> dense walk
        y = _random_decoded(1200, 1500, 1, 5)
    else:
        y = _random_decoded(900, 2, 3, 6)
    y = make_scores_unique(y, y.shape[1] - 4, kw["conf"])
    want, want_idx = nms_oracle(y, kw["conf"], kw["iou"], max_det=kw["max_det"])
    got, got_idx = non_max_suppression(y.to(cuda_device), kw["conf"], kw["iou"], max_det=kw["max_det"], return_idxs=True)
    assert sum(w.shape[0] for w in want) > 0
    assert_rows_equal(got, got_idx, want, want_idx, case)


def test_labels_and_end2end_shortcuts(cuda_device):
    from ultralytics_pro_b200.nms import non_max_suppression

    # a-priori labels (autolabelling, nms.py:100-105): rows appended after the candidates
    y = _random_decoded(400, 5, 2, 11)
    labels = [torch.tensor([[2.0, 100.0, 120.0, 30.0, 40.0], [4.0, 300.0, 310.0, 50.0, 20.0]]), torch.zeros((0, 5))]
    want, _ = nms_oracle(y, 0.3, 0.5, labels=labels)
    got = non_max_suppression(y.to(cuda_device), 0.3, 0.5, labels=[l.to(cuda_device) for l in labels])
    assert_rows_equal(got, None, want, None, "labels")
    # end-to-end layout (B, N, 6): threshold, cap, class filter (nms.py:66-70)
    e = torch.zeros(2, 50, 6)
    e[..., 4] = torch.rand(2, 50, generator=torch.Generator().manual_seed(3))
    e[..., 5] = torch.randint(0, 4, (2, 50), generator=torch.Generator().manual_seed(4)).float()
    e[..., :4] = torch.rand(2, 50, 4, generator=torch.Generator().manual_seed(5)) * 100
    for kwargs in (dict(max_det=7), dict(max_det=300, classes=[1, 3])):
        want, _ = nms_oracle(e, 0.4, 0.5, **kwargs)
        got = non_max_suppression(e.to(cuda_device), 0.4, 0.5, **kwargs)
        assert_rows_equal(got, None, want, None, f"end2end {kwargs}")


def test_head_post_processor_and_graph(cuda_device):
    """The cached-plan API and its CUDA-graph replay give the same rows as the functional call."""
    from ultralytics_pro_b200.head import postprocess_from_head
    from ultralytics_pro_b200.pipeline import HeadPostProcessor

    cfg = CONFIGS["c2_v8x_640_b64"]
    levels, _ = make_head_batch(cfg, batch=3, seed=51)
    dl = [lv.to(cuda_device) for lv in levels]
    want, want_idx = postprocess_from_head(dl, cfg.strides, cfg.nc, cfg.conf, cfg.iou, return_idxs=True)
    post = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou)
    got, got_idx = post(dl, return_idxs=True)
    assert_rows_equal(got, got_idx, [w.cpu() for w in want], [w.cpu() for w in want_idx], "post processor")
    graph = post.capture(dl)
    post.last.rows.zero_()
    graph.replay()
    got, got_idx = post.results(return_idxs=True)
    assert_rows_equal(got, got_idx, [w.cpu() for w in want], [w.cpu() for w in want_idx], "graph replay")


@pytest.mark.parametrize("batch", [1, 5])
def test_graphed_post_processor_host_counts(cuda_device, batch):
    """use_graph=True: the per-image counts come from the plan's mapped host buffer, written by the suppression kernel itself
    (ypb_nms_out.count_host) - no copy node; a single image skips the packing kernel and returns views of the plan's rows.
    Replays on changed input contents must follow the contents."""
    from ultralytics_pro_b200.head import postprocess_from_head
    from ultralytics_pro_b200.pipeline import HeadPostProcessor

    cfg = CONFIGS["c2_v8x_640_b64"]
    post = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, use_graph=True)
    static = [lv.to(cuda_device) for lv in make_head_batch(cfg, batch=batch, seed=60)[0]]
    for seed in (61, 62, 63):
        fresh = [lv.to(cuda_device) for lv in make_head_batch(cfg, batch=batch, seed=seed)[0]]
        for dst, src in zip(static, fresh):
            dst.copy_(src)
        want, want_idx = postprocess_from_head(fresh, cfg.strides, cfg.nc, cfg.conf, cfg.iou, return_idxs=True)
        got, got_idx = post(static, return_idxs=True)
        assert post.last.count_host is not None
        assert post.last.count_host.tolist() == [int(w.shape[0]) for w in want]
        assert_rows_equal([g.cpu() for g in got], [g.cpu() for g in got_idx], [w.cpu() for w in want], [w.cpu() for w in want_idx],
                          f"graph + host counts, seed {seed}")
    assert len(post._graphs) == 1


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("mode", ["obb_nc15", "xyxy", "nc20_tail"])
def test_decode_dense16_modes(cuda_device, dtype, mode):
    """The 16-bit dense decode kernel (16-row batches): rotated decode with the class rows entirely in the tail loop (nc = 15),
    the corner form, and a class count with whole batches AND a tail (nc = 20), against fp32 math on the same rounded inputs."""
    from ultralytics_pro_b200.head import decode_head
    from ultralytics_pro_b200.synth import HeadConfig

    if mode == "obb_nc15":
        cfg = CONFIGS["c5_obb_1024_b16"]
        levels, ang = make_head_batch(cfg, batch=2, seed=71, dtype=dtype)
        exact = obb_forward_oracle([lv.float() for lv in levels], ang.float(), cfg.strides, cfg.nc)
        dl, da = _to(cuda_device, levels, ang)
        got = decode_head(dl, cfg.strides, cfg.nc, angle=da, angle_is_logit=True, append_angle=True).float().cpu()
    elif mode == "xyxy":
        cfg = CONFIGS["c1_v8n_640_b1"]
        levels, _ = make_head_batch(cfg, batch=2, seed=72, dtype=dtype)
        exact = decode_oracle([lv.float() for lv in levels], cfg.strides, cfg.nc, xyxy=True)
        got = decode_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc, xyxy=True).float().cpu()
    else:
        cfg = HeadConfig("nc20", 320, (8, 16, 32), 20, 2, objects=5)
        levels, _ = make_head_batch(cfg, batch=2, seed=73, dtype=dtype)
        exact = decode_oracle([lv.float() for lv in levels], cfg.strides, cfg.nc)
        got = decode_head(_to(cuda_device, levels)[0], cfg.strides, cfg.nc).float().cpu()
    assert got.shape == exact.shape
    tol = 1e-2 * exact.abs() + 1e-2
    if mode == "obb_nc15":
        # the angle is rounded to 16 bits before the rotation (as the reference's tensor is): a centre moves by up to
        # |offset| * 2^-9 * |theta| pixels against fp32 math - allow it on the two centre rows, and require the aggregate error to be
        # no worse than the reference's own per-op-rounded chain
        tol[:, :2] += 1.5
        ref = obb_forward_oracle(levels, ang, cfg.strides, cfg.nc).float()
        assert float((got - exact).abs().mean()) <= float((ref - exact).abs().mean()) * 1.05 + 1e-6
    err = (got - exact).abs()
    assert not bool((err > tol).any()), f"max err {float(err.max())} at {int(err.argmax())}"
