"""Size-independent properties of the path at BASELINE.json's full configurations (C2 B=64, C5 B=16, C3 B=8), in addition to
the direct oracle comparisons of test_gpu_parity.py: what any correct non_max_suppression result must satisfy whatever the
input - checked on the CUDA results alone, so the sizes are not limited by what the CPU oracle finishes in seconds.

  * rows of an image are sorted by descending score, counts <= max_det, kept anchor indices unique and in range (nms.py:137-161);
  * greedy rule (nms.py:239-296): no two kept rows overlap by more than the threshold on the class-offset boxes the
    suppression ran on (nms.py:143-149) - evaluated here with plain torch arithmetic, not with the library;
  * idempotence: TorchNMS.nms / fast_nms over the kept rows of an image keeps every one of them, in order;
  * equivariance: reversing the image order of the batch reverses the result list, bit for bit (no cross-image state in the
    GPU-wide work lists, counters and lanes).
"""
import pytest
import torch

from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

pytestmark = pytest.mark.gpu


def _run(cfg, levels, ang):
    from ultralytics_pro_b200.head import postprocess_from_head

    return postprocess_from_head(levels, cfg.strides, cfg.nc, cfg.conf, cfg.iou, multi_label=cfg.multi_label,
                                 agnostic=cfg.agnostic, max_det=cfg.max_det, max_nms=cfg.max_nms, angle_logits=ang,
                                 return_idxs=True)


def _check_basic(cfg, rows, idx):
    assert len(rows) == len(idx)
    total = 0
    for b, (r, i) in enumerate(zip(rows, idx)):
        n = r.shape[0]
        total += n
        assert n <= cfg.max_det and i.numel() == n, f"image {b}"
        if n == 0:
            continue
        s = r[:, 4]
        assert bool((s[1:] <= s[:-1]).all()), f"image {b}: scores not sorted"
        assert bool((s > cfg.conf).all()), f"image {b}: a kept score is not above conf"
        iv = i.view(-1)
        assert int(iv.min()) >= 0 and int(iv.max()) < cfg.anchors, f"image {b}: anchor index out of range"
        if not cfg.multi_label:
            assert torch.unique(iv).numel() == n, f"image {b}: an anchor was kept twice"
        else:  # one row per (anchor, class) pair
            pair = iv * cfg.nc + r[:, 5].long()
            assert torch.unique(pair).numel() == n, f"image {b}: an (anchor, class) pair was kept twice"
        c = r[:, 5]
        assert bool(((c >= 0) & (c < cfg.nc) & (c == c.round())).all()), f"image {b}: class column"
    assert total > 0
    return total


def _offset_iou(r, max_wh):
    """Pairwise IoU of the kept rows on the boxes the greedy walk saw: box + cls * max_wh (nms.py:143-149), fp32, the
    reference's operation order (nms.py:281-292)."""
    b = r[:, :4] + r[:, 5:6] * max_wh
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    w = (torch.minimum(b[:, None, 2], b[None, :, 2]) - torch.maximum(b[:, None, 0], b[None, :, 0])).clamp_(min=0)
    h = (torch.minimum(b[:, None, 3], b[None, :, 3]) - torch.maximum(b[:, None, 1], b[None, :, 1])).clamp_(min=0)
    inter = w * h
    return inter / (area[:, None] + area[None, :] - inter)


@pytest.mark.parametrize("name,batch", [("c2_v8x_640_b64", 64), ("c4_p6_1280_b16", 16), ("c3_val_stress_b32", 8)])
def test_greedy_result_properties_at_full_size(cuda_device, name, batch):
    from ultralytics_pro_b200.nms import TorchNMS

    cfg = CONFIGS[name]
    levels, ang = make_head_batch(cfg, batch=batch, seed=71, device=cuda_device)
    rows, idx = _run(cfg, levels, ang)
    _check_basic(cfg, rows, idx)
    max_wh = 0.0 if cfg.agnostic else 7680.0
    thr = torch.tensor(cfg.iou, dtype=torch.float64).item()
    for b, r in enumerate(rows):
        n = r.shape[0]
        if n < 2:
            continue
        iou = _offset_iou(r, max_wh)
        upper = torch.triu(iou, diagonal=1)
        # kept pairs never exceed the threshold (the walk suppresses iff fp32 IoU, widened to double, > thr)
        assert float(upper.max()) <= thr, f"image {b}: two kept rows overlap by {float(upper.max())}"
        # idempotence: a second NMS over the kept rows keeps them all, in the same order
        again = TorchNMS.nms(r[:, :4] + r[:, 5:6] * max_wh, r[:, 4], cfg.iou)
        assert torch.equal(again, torch.arange(n, device=r.device)), f"image {b}: NMS of the kept rows is not the identity"


def test_rotated_result_properties_at_full_size(cuda_device):
    from ultralytics_pro_b200.nms import TorchNMS, batch_probiou

    cfg = CONFIGS["c5_obb_1024_b16"]
    levels, ang = make_head_batch(cfg, batch=16, seed=72, device=cuda_device)
    rows, idx = _run(cfg, levels, ang)
    _check_basic(cfg, rows, idx)
    for b, r in enumerate(rows):
        n = r.shape[0]
        if n < 2:
            continue
        assert r.shape[1] == 7
        # Fast-NMS (nms.py:217-223): a kept row has no higher-ranked row AT ALL with ProbIoU >= thr - in particular no kept one:
        # Fast-NMS over the kept rows (class offset on the centres, nms.py:146) is the identity
        boxes = torch.cat((r[:, :2] + r[:, 5:6] * 7680.0, r[:, 2:4], r[:, 6:7]), 1)
        again = TorchNMS.fast_nms(boxes, r[:, 4], cfg.iou, iou_func=batch_probiou)
        assert torch.equal(again, torch.arange(n, device=r.device)), f"image {b}: Fast-NMS of the kept rows is not the identity"


@pytest.mark.parametrize("name,batch", [("c2_v8x_640_b64", 64), ("c5_obb_1024_b16", 16)])
def test_batch_order_equivariance_at_full_size(cuda_device, name, batch):
    cfg = CONFIGS[name]
    levels, ang = make_head_batch(cfg, batch=batch, seed=73, device=cuda_device)
    rows, idx = _run(cfg, levels, ang)
    flipped = [lv.flip(0).contiguous() for lv in levels]
    fang = ang.flip(0).contiguous() if ang is not None else None
    rrows, ridx = _run(cfg, flipped, fang)
    assert sum(r.shape[0] for r in rows) > 0
    for b in range(batch):
        assert torch.equal(rows[b], rrows[batch - 1 - b]), f"image {b}: rows depend on the position in the batch"
        assert torch.equal(idx[b].view(-1), ridx[batch - 1 - b].view(-1)), f"image {b}: kept anchors depend on the position in the batch"
