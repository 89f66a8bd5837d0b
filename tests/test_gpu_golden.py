"""GPU parity against the committed golden vectors of the LIVE reference (tests/golden/*.npz, oracle/make_golden.py):
decode within the north-star tolerance, NMS rows / kept indices bit-exact, through the C-ABI."""
import glob
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_decode_against_reference_golden(cuda_device, path):
    from ultralytics_pro_b200.head import decode_head

    z, meta = _load(path)
    levels = [torch.from_numpy(z[f"level{i}"]).to(cuda_device) for i in range(len(meta["strides"]))]
    if meta["rotated"]:
        ang = torch.from_numpy(z["angle_logits"]).to(cuda_device)
        got = decode_head(levels, meta["strides"], meta["nc"], meta["reg_max"], angle=ang, angle_is_logit=True, append_angle=True)
    else:
        got = decode_head(levels, meta["strides"], meta["nc"], meta["reg_max"])
    want = torch.from_numpy(z["decoded_raw"])
    got = got.cpu()
    assert got.shape == want.shape
    tol = 1e-5 * want.abs() + 1e-5 * meta["imgsz"]
    assert not bool(((got - want).abs() > tol).any()), f"max abs diff {float((got - want).abs().max())}"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_nms_against_reference_golden(cuda_device, path):
    from ultralytics_pro_b200.nms import non_max_suppression

    z, meta = _load(path)
    y = torch.from_numpy(z["decoded"]).to(cuda_device)
    for ci, kw in enumerate(meta["calls"]):
        out, keep = non_max_suppression(y.clone(), nc=meta["nc"], return_idxs=True, **kw)
        for b in range(meta["batch"]):
            want_rows = torch.from_numpy(z[f"call{ci}_rows{b}"])
            want_idx = torch.from_numpy(z[f"call{ci}_idx{b}"])
            assert torch.equal(keep[b].cpu(), want_idx), f"{meta['name']} call {ci} {kw} image {b}: kept indices differ"
            assert torch.equal(out[b].cpu(), want_rows), f"{meta['name']} call {ci} {kw} image {b}: rows differ"


def test_input_not_mutated_by_default_and_mutation_switch(cuda_device):
    import ultralytics_pro_b200.nms as nms

    z, meta = _load(GOLDEN[0])
    y = torch.from_numpy(z["decoded"]).to(cuda_device)
    before = y.clone()
    nms.non_max_suppression(y, 0.25, 0.7, nc=meta["nc"])
    assert torch.equal(y, before)
    nms.MUTATE_INPUT_LIKE_REFERENCE = True
    try:
        nms.non_max_suppression(y, 0.25, 0.7, nc=meta["nc"])
    finally:
        nms.MUTATE_INPUT_LIKE_REFERENCE = False
    want = before.clone()
    want[:, :2] = before[:, :2] - before[:, 2:4] / 2
    want[:, 2:4] = before[:, :2] + before[:, 2:4] / 2
    assert torch.equal(y, want)  # nms.py:86 side effect reproduced on request
