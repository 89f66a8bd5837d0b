"""Segment masks (part 2): process_mask / process_mask_native (utils/ops.py:489-541).

CPU: the oracle restatement against the live-reference golden masks (tests/golden/post/masks.npz) - identical.
GPU: the fused CUDA kernel (through the C-ABI) against the golden masks.  The result is a threshold of a float value
(32-term dot product, bilinear blend) whose summation order differs between MKL/ATen and the kernel, so a pixel may
differ only where the oracle's pre-threshold value is within 1e-4 of zero (values are O(1)); everything else is exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro

PATH = os.path.join(os.path.dirname(__file__), "golden", "post", "masks.npz")
Z = np.load(PATH)
META = json.loads(bytes(Z["meta"]).decode())
IDS = [m["name"] for m in META]
DT = {"float32": torch.float32, "float16": torch.float16}
TOL = 1e-4


def _case(i):
    m = META[i]
    protos = torch.from_numpy(Z[f"m{i}_protos"]).to(DT[m["dtype"]])
    coef, boxes = torch.from_numpy(Z[f"m{i}_coef"]), torch.from_numpy(Z[f"m{i}_boxes"])
    n = int(np.prod(m["out_shape"]))
    want = np.unpackbits(Z[f"m{i}_out"])[:n].reshape(m["out_shape"]).astype(np.uint8)
    return m, protos, coef, boxes, want


def _oracle(m, protos, coef, boxes):
    if m["kind"] == "process_mask_native":
        return ro.process_mask_native_oracle(protos, coef, boxes, m["shape"])
    return ro.process_mask_oracle(protos, coef, boxes, m["shape"], upsample=m["kind"] == "process_mask_up")


@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_oracle_masks_match_reference_golden(i):
    m, protos, coef, boxes, want = _case(i)
    got, _ = _oracle(m, protos, coef, boxes)
    assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), want)


def _check(got, want, values, what):
    got = got.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.uint8, what
    bad = got != want
    if bad.any():
        worst = np.abs(values.numpy()[bad]).max()
        assert worst < TOL, f"{what}: {int(bad.sum())} pixels differ, one with |value| = {worst}"
    assert bad.mean() < 1e-4, f"{what}: {int(bad.sum())} of {bad.size} pixels differ"


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_cuda_masks_match_reference_golden(cuda_device, i):
    from ultralytics_pro_b200 import ops

    m, protos, coef, boxes, want = _case(i)
    _, values = _oracle(m, protos, coef, boxes)
    p, c, b = protos.to(cuda_device), coef.to(cuda_device), boxes.to(cuda_device)
    if m["kind"] == "process_mask_native":
        got = ops.process_mask_native(p, c, b, m["shape"])
    else:
        got = ops.process_mask(p, c, b, m["shape"], upsample=m["kind"] == "process_mask_up")
    _check(got, want, values, m["name"])
    # rows of an NMS result: coefficients / boxes as strided column views of one (n, 38) tensor
    rows = torch.cat([b, torch.zeros(len(b), 2, device=cuda_device), c], 1)
    if m["kind"] == "process_mask_up":
        _check(ops.process_mask(p, rows[:, 6:], rows[:, :4], m["shape"], upsample=True), want, values, m["name"] + " (views)")


@pytest.mark.gpu
def test_batched_masks_equal_per_image_calls(cuda_device):
    """One launch for the whole batch (packed output, device prefix of the counts) == process_mask image by image."""
    from ultralytics_pro_b200 import ops

    g = torch.Generator().manual_seed(9)
    B, M, C, mh, mw = 4, 12, 32, 40, 40
    protos = torch.randn(B, C, mh, mw, generator=g).to(cuda_device)
    rows = torch.zeros(B, M, 6 + C)
    xy = torch.rand(B, M, 2, generator=g) * 100
    rows[..., :2], rows[..., 2:4] = xy, xy + torch.rand(B, M, 2, generator=g) * 60 + 1
    rows[..., 6:] = torch.randn(B, M, C, generator=g)
    rows = rows.to(cuda_device)
    counts = [5, 0, 12, 1]
    outs = ops.process_masks_batched(protos, rows, counts, (160, 160), upsample=True)
    assert [o.shape[0] for o in outs] == counts
    for b, n in enumerate(counts):
        one = ops.process_mask(protos[b], rows[b, :n, 6:], rows[b, :n, :4], (160, 160), upsample=True)
        assert torch.equal(outs[b], one), f"image {b}"
    # odd output width (not a multiple of 16): scalar store path
    odd = ops.process_mask(protos[0], rows[0, :5, 6:], rows[0, :5, :4], (150, 157), upsample=True)
    _, val = ro.process_mask_oracle(protos[0].cpu(), rows[0, :5, 6:].cpu(), rows[0, :5, :4].cpu(), (150, 157), upsample=True)
    _check(odd, (val > 0).numpy().astype(np.uint8), val, "odd width")
    assert ops.process_mask(protos[0], rows[0, :0, 6:], rows[0, :0, :4], (160, 160), upsample=True).shape == (0, 160, 160)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["process_mask_up", "process_mask_native", "process_mask"])
def test_two_step_masks_equal_single_kernel(cuda_device, kind, monkeypatch):
    """The memset + work-list + compute-listed-tiles form (used for results >= ops.TWO_STEP_MIN_BYTES) writes exactly the
    bytes the single kernel writes: boxes from a few pixels to the whole image, batched, odd sizes, repeated launches."""
    from ultralytics_pro_b200 import ops

    g = torch.Generator().manual_seed(17)
    B, M, C, mh, mw = 3, 40, 32, 48, 64
    protos = torch.randn(B, C, mh, mw, generator=g).to(cuda_device)
    rows = torch.zeros(B, M, 6 + C)
    xy = torch.rand(B, M, 2, generator=g) * torch.tensor([200.0, 150.0])
    size = torch.exp(torch.rand(B, M, 2, generator=g) * 5)            # 1 .. 150 px
    rows[..., :2], rows[..., 2:4] = xy, xy + size
    rows[0, 0, :4] = torch.tensor([0.0, 0.0, 256.0, 192.0])            # the whole image
    rows[..., 6:] = torch.randn(B, M, C, generator=g)
    rows = rows.to(cuda_device)
    shape = (192, 256) if kind != "process_mask_native" else (381, 509)

    def run():
        if kind == "process_mask_up":
            return ops.process_masks_batched(protos, rows, [M, 7, 0], shape, upsample=True)
        if kind == "process_mask_native":
            return [ops.process_mask_native(protos[0], rows[0, :, 6:], rows[0, :, :4], shape)]
        return [ops.process_mask(protos[1], rows[1, :, 6:], rows[1, :, :4], shape, upsample=False)]

    monkeypatch.setattr(ops, "TWO_STEP_MIN_BYTES", 1 << 60)
    single = run()
    monkeypatch.setattr(ops, "TWO_STEP_MIN_BYTES", 0)
    for _ in range(2):
        two = run()
        assert len(two) == len(single) and all(torch.equal(a, b) for a, b in zip(two, single))
    assert sum(int(t.sum()) for t in single) > 0


def test_mask_symbols_exported():
    from ultralytics_pro_b200 import _cabi

    assert "ypb_process_mask" in _cabi.EXPORTS
