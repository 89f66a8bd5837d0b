"""Validator matching: engine/validator.py:267-307 match_predictions + metrics.py:54 box_iou.

CPU: the sort-free oracle restatement against the live-reference golden (tests/golden/post/matching.npz) - identical.
GPU: the kernel (C-ABI) in matrix mode (drop-in for match_predictions), in boxes mode (drop-in for
DetectionValidator._process_batch, IoU computed on the fly) and batched over a validation batch - identical."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "post", "matching.npz"))
META = json.loads(bytes(Z["meta"]).decode())
IDS = [m["name"] for m in META]
IOUV = Z["iouv"]


@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_oracle_matching_matches_reference_golden(i):
    iou = ro.box_iou_oracle(Z[f"q{i}_gt"], Z[f"q{i}_pred"])
    assert np.array_equal(iou, Z[f"q{i}_iou"]), "box_iou restatement must be bit-identical"
    tp = ro.match_predictions_oracle(Z[f"q{i}_pcls"], Z[f"q{i}_gcls"], iou, IOUV)
    assert np.array_equal(tp, Z[f"q{i}_tp"])


def test_match_symbol_exported():
    from ultralytics_pro_b200 import _cabi

    assert "ypb_match_predictions" in _cabi.EXPORTS


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_cuda_matching_matches_reference_golden(cuda_device, i):
    from ultralytics_pro_b200 import val

    dev = cuda_device
    gt, gcls = torch.from_numpy(Z[f"q{i}_gt"]).to(dev), torch.from_numpy(Z[f"q{i}_gcls"]).to(dev)
    pred, pcls = torch.from_numpy(Z[f"q{i}_pred"]).to(dev), torch.from_numpy(Z[f"q{i}_pcls"]).to(dev)
    want = Z[f"q{i}_tp"]
    me = types.SimpleNamespace(iouv=torch.from_numpy(IOUV), niou=len(IOUV))
    got = val.match_predictions(me, pcls, gcls, torch.from_numpy(Z[f"q{i}_iou"]).to(dev))  # matrix mode
    assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), want)
    tp = val.process_batch(me, {"bboxes": pred, "cls": pcls}, {"bboxes": gt, "cls": gcls})["tp"]  # boxes mode
    assert tp.dtype == bool and np.array_equal(tp, want)
    with pytest.raises(NotImplementedError):
        val.match_predictions(me, pcls, gcls, torch.from_numpy(Z[f"q{i}_iou"]).to(dev), use_scipy=True)


@pytest.mark.gpu
def test_cuda_matching_batched_and_empty(cuda_device):
    from ultralytics_pro_b200 import val

    dev = cuda_device
    B, MD = len(META), 300
    rows = torch.zeros(B, MD, 6)
    counts, labels, lcounts = [], [], []
    for i in range(B):
        n = Z[f"q{i}_pred"].shape[0]
        rows[i, :n, :4] = torch.from_numpy(Z[f"q{i}_pred"])
        rows[i, :n, 5] = torch.from_numpy(Z[f"q{i}_pcls"])
        rows[i, n:, :4] = torch.tensor([0.0, 0.0, 600.0, 600.0])  # garbage past the count must not match anything
        counts.append(n)
        labels.append(np.concatenate([Z[f"q{i}_gcls"][:, None], Z[f"q{i}_gt"]], 1))
        lcounts.append(len(Z[f"q{i}_gt"]))
    out = val.match_batch(IOUV.tolist(), rows.to(dev), torch.tensor(counts, dtype=torch.int32, device=dev),
                          torch.from_numpy(np.concatenate(labels)).to(dev), lcounts).cpu().numpy()
    for i in range(B):
        n = counts[i]
        assert np.array_equal(out[i, :n].astype(bool), Z[f"q{i}_tp"]), META[i]["name"]
        assert not out[i, n:].any()
    me = types.SimpleNamespace(iouv=torch.from_numpy(IOUV), niou=len(IOUV))
    e = val.process_batch(me, {"bboxes": torch.zeros(0, 4, device=dev), "cls": torch.zeros(0, device=dev)},
                          {"bboxes": torch.zeros(3, 4, device=dev), "cls": torch.zeros(3, device=dev)})["tp"]
    assert e.shape == (0, len(IOUV))
    z = val.match_iou_matrix(IOUV.tolist(), torch.zeros(4, device=dev), torch.zeros(0, device=dev), torch.zeros(0, 4, device=dev))
    assert z.shape == (4, len(IOUV)) and not z.any()
