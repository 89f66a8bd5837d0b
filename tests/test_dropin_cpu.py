"""CPU suite, part 3: the drop-in boundary itself - the lazily decoded tensor that links ``Detect._inference`` to
``non_max_suppression`` (ultralytics_pro_b200/lazy.py), the host-stride cache, and the rebinding of the reference's own
classes.  No kernel is launched: the lazy tensor's materialiser is replaced by the CPU oracle decode (test
infrastructure), which is enough to pin the dispatch mechanics every reference call site relies on."""
import pytest
import torch

from oracle.postproc_oracle import decode_oracle, nms_oracle
from tests.helpers import small_cfg
from ultralytics_pro_b200 import lazy
from ultralytics_pro_b200.synth import make_head_batch


@pytest.fixture()
def oracle_materializer(monkeypatch):
    calls = []

    def mat(rec):
        calls.append(rec)
        return decode_oracle(rec.levels, rec.strides, rec.nc, rec.reg_max, xyxy=rec.xyxy)

    monkeypatch.setattr(lazy, "_materializer", mat)
    return calls


def _lazy_and_dense(cfg, seed=3, xyxy=False):
    levels, _ = make_head_batch(cfg, seed=seed)
    dense = decode_oracle(levels, cfg.strides, cfg.nc, cfg.reg_max, xyxy=xyxy)
    return lazy.LazyDecoded.from_head(levels, cfg.strides, cfg.nc, cfg.reg_max, xyxy), dense, levels


def test_lazy_tensor_is_a_tensor_with_the_dense_metadata_and_no_work(oracle_materializer):
    cfg = small_cfg(batch=2)
    y, dense, _ = _lazy_and_dense(cfg)
    assert isinstance(y, torch.Tensor)
    assert y.shape == dense.shape and y.dtype == dense.dtype and y.device == dense.device and y.dim() == 3
    assert y.shape[-1] != 6 and y.is_contiguous() and y.stride() == dense.stride()
    assert "pending" in repr(y)
    assert oracle_materializer == [] and y.head_record() is not None  # nothing above needed the values


@pytest.mark.parametrize("inference_mode", [False, True])
def test_every_kind_of_consumer_sees_the_dense_values(oracle_materializer, inference_mode):
    cfg = small_cfg(batch=2)
    ctx = torch.inference_mode() if inference_mode else torch.no_grad()
    with ctx:
        consumers = {
            "permute (Detect.forward_end2end, head.py:148)": lambda t: t.permute(0, 2, 1),
            "cat (Segment/Pose/OBB.forward, head.py:837)": lambda t: torch.cat([t, t[:, :2]], 1),
            "index": lambda t: t[1, 4:, ::7],
            "amax (nms.py:76)": lambda t: t[:, 4:].amax(1) > 0.25,
            "transpose (nms.py:84)": lambda t: t.transpose(-1, -2),
            "to half": lambda t: t.to(torch.float16),
            "clone": lambda t: t.clone(),
            "arithmetic": lambda t: t * 2 + 1,
            "sum": lambda t: t.sum(),
        }
        for what, fn in consumers.items():
            y, dense, _ = _lazy_and_dense(cfg)
            got, want = fn(y), fn(dense)
            assert type(got) is torch.Tensor, what
            assert torch.equal(got, want), what
            assert y.head_record() is None and "materialized" in repr(y), what
        assert len(oracle_materializer) == len(consumers)
        # decoded once, however many consumers follow
        y, dense, _ = _lazy_and_dense(cfg)
        n0 = len(oracle_materializer)
        assert torch.equal(y + 1, dense + 1) and torch.equal(y[:, :4], dense[:, :4]) and torch.equal(y.cpu(), dense)
        assert len(oracle_materializer) == n0 + 1


def test_in_place_writes_land_in_the_materialised_tensor(oracle_materializer):
    """nms.py:86 rewrites the box channels of its input in place; a caller that does the same to the lazy tensor must
    read its own writes back."""
    cfg = small_cfg(batch=1)
    y, dense, _ = _lazy_and_dense(cfg)
    y[:, :2] = 0.0
    dense[:, :2] = 0.0
    assert torch.equal(y.clone(), dense) and len(oracle_materializer) == 1


def test_reference_nms_accepts_the_lazy_tensor(oracle_materializer):
    """The oracle restatement of non_max_suppression (same torch operators as nms.py:58-166) run on the lazy tensor gives
    what it gives on the dense tensor: every operator it uses dispatches through the wrapper."""
    cfg = small_cfg(batch=2)
    y, dense, _ = _lazy_and_dense(cfg, seed=9)
    want, want_idx = nms_oracle(dense.clone(), 0.25, 0.7, nc=cfg.nc)
    got, got_idx = nms_oracle(y, 0.25, 0.7, nc=cfg.nc)
    for g, w, gi, wi in zip(got, want, got_idx, want_idx):
        assert torch.equal(g, w) and torch.equal(gi, wi)
    assert sum(len(w) for w in want) > 0


def test_live_reference_module_returns_lazy_only_for_plain_detect_heads(oracle_materializer, monkeypatch):
    """With the reference mounted: install() rebinds _inference / decode_bboxes / DFL.forward on the reference's OWN classes,
    CPU tensors still run the untouched reference code, and the attribute reads of detect_inference work on a real
    ``Detect`` / ``OBB`` / ``Pose`` module (head.py:70-93)."""
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    ref = load_reference()
    import ultralytics_pro_b200.head as our_head
    import ultralytics_pro_b200.patch as patch

    done = patch.install()
    try:
        for cname in ("Detect", "MAFDetect", "IDetect", "DDetect"):
            assert f"ultralytics.nn.modules.head.{cname}._inference" in done
            assert f"ultralytics.nn.modules.head.{cname}.decode_bboxes" in done
        for cname in ("OBB", "MAFOBB", "IOBB", "DOBB"):
            assert f"ultralytics.nn.modules.head.{cname}.decode_bboxes" in done
        assert "ultralytics.nn.modules.block.DFL.forward" in done
        assert lazy.ENABLED
        torch.manual_seed(0)
        det = ref.Detect(nc=5, ch=(16, 32)).eval()
        det.stride = torch.tensor([8.0, 16.0])
        feats = [torch.randn(2, 16, 8, 8), torch.randn(2, 32, 4, 4)]
        with torch.inference_mode():
            y_patched, raw = det([f.clone() for f in feats])          # CPU: the reference's own code runs
        patch.uninstall()
        with torch.inference_mode():
            y_ref, _ = det([f.clone() for f in feats])
        assert type(y_patched) is torch.Tensor and torch.equal(y_patched, y_ref)
        # the drop-in itself on the real module (attribute surface), kernels replaced by the oracle decode
        patch.install()
        monkeypatch.setattr(our_head, "decode_head", lambda lv, st, nc, reg_max=16, **kw: decode_oracle(lv, st, nc, reg_max, xyxy=kw.get("xyxy", False)))
        det.shape = None
        with torch.inference_mode():
            y_lazy = our_head.detect_inference(det, raw)
            assert isinstance(y_lazy, lazy.LazyDecoded) and y_lazy.shape == y_ref.shape
            assert our_head.host_strides(det) == (8.0, 16.0)
            assert torch.equal(det.anchors, _ref_anchor_cache(ref, raw, det.stride)[0])
            assert torch.equal(det.strides, _ref_anchor_cache(ref, raw, det.stride)[1])
            assert float((y_lazy - y_ref).abs().max()) < 1e-4
            # heads that concatenate their own channels right away get the dense tensor, not the wrapper
            obb = ref.OBB(nc=3, ne=1, ch=(16, 32)).eval()
            obb.stride = torch.tensor([8.0, 16.0])
            assert our_head._has_riders(obb) and not our_head._has_riders(det)
    finally:
        patch.uninstall()
    assert not lazy.ENABLED


def _ref_anchor_cache(ref, feats, stride):
    a, s = ref.tal.make_anchors(feats, stride, 0.5)
    return a.transpose(0, 1), s.transpose(0, 1)


def test_host_strides_are_read_back_once_per_stride_tensor():
    import types

    from ultralytics_pro_b200.head import host_strides

    class CountingTensor(torch.Tensor):
        reads = 0

        def tolist(self):
            CountingTensor.reads += 1
            return super().tolist()

    m = types.SimpleNamespace()
    m.stride = torch.tensor([8.0, 16.0, 32.0]).as_subclass(CountingTensor)
    for _ in range(5):
        assert host_strides(m) == (8.0, 16.0, 32.0)
    assert CountingTensor.reads == 1
    m.stride = torch.tensor([4.0, 8.0]).as_subclass(CountingTensor)  # BaseModel._apply replaces the tensor (tasks.py:1193-1210)
    assert host_strides(m) == (4.0, 8.0) and CountingTensor.reads == 2
