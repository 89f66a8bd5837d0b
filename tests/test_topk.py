"""End2end top-k (second half: nn/modules/head.py:193-214 Detect.postprocess).

CPU: the single-ranking oracle restatement against the live-reference golden (tests/golden/post/topk.npz) - identical.
GPU: detect_postprocess (two passes of the filter + sort kernels through the C-ABI) - identical."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import result_ops_oracle as ro

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "post", "topk.npz"))
META = json.loads(bytes(Z["meta"]).decode())
IDS = [m["name"] for m in META]


@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_oracle_topk_matches_reference_golden(i):
    m = META[i]
    got = ro.detect_postprocess_oracle(torch.from_numpy(Z[f"t{i}_preds"]), m["max_det"], m["nc"])
    assert np.array_equal(got.numpy(), Z[f"t{i}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)), ids=IDS)
def test_cuda_topk_matches_reference_golden(cuda_device, i):
    from ultralytics_pro_b200.head import detect_postprocess

    m = META[i]
    preds = torch.from_numpy(Z[f"t{i}_preds"]).to(cuda_device)
    got = detect_postprocess(preds, m["max_det"], m["nc"])
    assert np.array_equal(got.cpu().numpy(), Z[f"t{i}_out"]), m["name"]
    # the layout v10Detect produces: a permuted view of (B, 4+nc, A)
    view = preds.transpose(1, 2).contiguous().permute(0, 2, 1)
    assert np.array_equal(detect_postprocess(view, m["max_det"], m["nc"]).cpu().numpy(), Z[f"t{i}_out"])
