"""The exporter's embedded NMS (engine/exporter.py:1389-1481 NMSModel.forward after the model call).

CPU: the oracle restatement against the live-reference golden (tests/golden/post/nms_model.npz) - identical.
GPU: nms_model_postprocess (C-ABI, ypb_nms_from_dense with the normalised-offset mode) - identical rows, zero padding
included, and no host synchronisation (the result is one fixed-size tensor)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "post", "nms_model.npz"))
META = json.loads(bytes(Z["meta"]).decode())
IDS = [m["name"] for m in META]


def _kw(m):
    return dict(conf=m["conf"], iou=m["iou"], max_det=m["max_det"], agnostic_nms=m["agnostic_nms"])


@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_oracle_nms_model_matches_reference_golden(i):
    m = META[i]
    got = ro.nms_model_oracle(torch.from_numpy(Z[f"n{i}_pred"]), (m["imgsz"], m["imgsz"]), m["nc"], **_kw(m))
    assert np.array_equal(got.numpy(), Z[f"n{i}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_cuda_nms_model_matches_reference_golden(cuda_device, i):
    from ultralytics_pro_b200.export_nms import nms_model_postprocess

    m = META[i]
    pred = torch.from_numpy(Z[f"n{i}_pred"]).to(cuda_device)
    out, cnt = nms_model_postprocess(pred, (m["imgsz"], m["imgsz"]), m["nc"], return_count=True, **_kw(m))
    want = Z[f"n{i}_out"]
    assert tuple(out.shape) == want.shape
    assert np.array_equal(out.cpu().numpy(), want), m["name"]
    assert cnt.cpu().tolist() == [int((want[b, :, 4] > 0).sum()) for b in range(want.shape[0])]
    # the ordinary flavour must be untouched by the new parameters
    from ultralytics_pro_b200.nms import non_max_suppression
    from oracle.postproc_oracle import nms_oracle

    xywh = pred.clone()
    xywh[:, 0], xywh[:, 1] = (pred[:, 0] + pred[:, 2]) / 2, (pred[:, 1] + pred[:, 3]) / 2
    xywh[:, 2], xywh[:, 3] = pred[:, 2] - pred[:, 0], pred[:, 3] - pred[:, 1]
    got, idx = non_max_suppression(xywh, m["conf"], m["iou"], nc=m["nc"], return_idxs=True)
    ref, ridx = nms_oracle(xywh.cpu(), m["conf"], m["iou"], nc=m["nc"])
    for b in range(len(ref)):
        assert torch.equal(idx[b].cpu(), ridx[b].view(-1)) and torch.equal(got[b].cpu(), ref[b])
