"""CPU suite, part 2: host-side logic - the C-ABI surface, threshold arithmetic, synthetic data, sharding and the
two-rank gather (gloo), the reference patch points.  No kernel is launched here."""
import ctypes
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from ultralytics_pro_b200 import _cabi, build

    path = build.build_library()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "yolopost_b200.h")).read()
    declared = set(re.findall(r"YPB_API\s+[\w\s\*]+?\b(ypb_\w+)\s*\(", header))
    assert declared, "no YPB_API declarations parsed"
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/yolopost_b200.h but not exported"
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib.ypb_abi_version.restype = ctypes.c_int
    assert lib.ypb_abi_version() == _cabi.ABI_VERSION


def test_host_only_entry_points():
    """Entry points that never touch the device: workspace sizing and argument validation."""
    from ultralytics_pro_b200 import _cabi

    lib = _cabi.load()
    small = lib.ypb_nms_workspace_bytes(1, 8400, 8400, 300, 30000, 0)
    big = lib.ypb_nms_workspace_bytes(64, 8400, 8400, 300, 30000, 0)
    assert 0 < small < big
    assert lib.ypb_nms_workspace_bytes(32, 8400, 8400 * 80, 300, 30000, 0) > big
    assert lib.ypb_nms_boxes_workspace_bytes(1000) > 0
    # invalid descriptors are rejected before any launch
    d = _cabi.HeadDesc()
    d.num_levels = 0
    rc = lib.ypb_decode_dense(ctypes.byref(d), None, 0, 0, 0, None, 0, 0, 0, None)
    assert rc == -1 and b"num_levels" in lib.ypb_last_error_string()
    d.num_levels, d.batch, d.nc, d.reg_max, d.dtype = 1, 1, 80, 8, 0
    rc = lib.ypb_decode_dense(ctypes.byref(d), None, 0, 0, 0, None, 0, 0, 0, None)
    assert rc == -2 and b"reg_max" in lib.ypb_last_error_string()


def test_cpu_tensors_are_rejected_loudly():
    from ultralytics_pro_b200.head import decode_head
    from ultralytics_pro_b200.nms import TorchNMS, non_max_suppression

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        non_max_suppression(torch.zeros(1, 84, 10), 0.25, 0.7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        decode_head([torch.zeros(1, 144, 4, 4)], (8,), 80)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TorchNMS.nms(torch.zeros(2, 4), torch.zeros(2), 0.5)
    with pytest.raises(AssertionError):
        non_max_suppression(torch.zeros(1, 84, 10), -0.1, 0.7)


def test_threshold_arithmetic():
    from ultralytics_pro_b200 import _cabi

    f32 = lambda v: struct.unpack("f", struct.pack("f", v))[0]
    for thr in (0.0, 0.25, 0.45, 0.5, 0.6, 0.7, 0.001, 1.0, 0.3333333):
        eff = _cabi.largest_f32_not_above(thr)
        assert eff <= thr and f32(eff) == eff
        nxt = np.nextafter(np.float32(eff), np.float32(2.0))
        assert float(nxt) > thr
        # "x > thr in double" == "x > eff in float" on both neighbours
        for x in (np.float32(eff), nxt):
            assert (float(x) > thr) == (x > np.float32(eff))
    assert _cabi.largest_f32_not_above(0.6) < f32(0.6)      # float32(0.6) rounds up
    assert _cabi.largest_f32_not_above(0.7) == f32(0.7)      # float32(0.7) rounds down
    assert _cabi.round_to_dtype(0.3, torch.bfloat16) == 0.30078125
    assert _cabi.round_to_dtype(0.25, torch.float16) == 0.25
    # the same cast torch applies to the scalar of `tensor > conf`
    assert bool((torch.tensor([0.30078125], dtype=torch.bfloat16) > 0.3).item()) is False


def test_synthetic_generator_is_deterministic_and_shardable():
    from ultralytics_pro_b200.synth import CONFIGS, HeadConfig, make_head_batch

    cfg = HeadConfig("t", 160, (8, 16, 32), 80, 4, objects=5)
    a, _ = make_head_batch(cfg, seed=7)
    b, _ = make_head_batch(cfg, seed=7)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    lo, _ = make_head_batch(cfg, batch=2, seed=7, first_image=2)
    assert all(torch.equal(x[2:], y) for x, y in zip(a, lo))
    assert a[0].shape == (4, 144, 20, 20) and cfg.anchors == 525
    assert CONFIGS["c2_v8x_640_b64"].anchors == 8400 and CONFIGS["c4_p6_1280_b16"].anchors == 34000
    assert CONFIGS["c5_obb_1024_b16"].anchors == 21504
    obb = HeadConfig("o", 128, (8, 16, 32), 15, 2, rotated=True, objects=4)
    lv, ang = make_head_batch(obb, seed=1)
    assert lv[0].shape[1] == 79 and ang.shape == (2, 1, obb.anchors)


def test_shard_range_covers_batch():
    from ultralytics_pro_b200.dist import shard_range

    for batch in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_pack_unpack_roundtrip():
    from ultralytics_pro_b200.dist import pack_results, unpack_results

    rows = torch.randn(3, 5, 7)
    count = torch.tensor([5, 0, 2], dtype=torch.int32)
    packed = pack_results(rows, count, pad_batch=4)
    assert packed.shape == (4, 1 + 35)
    r2, c2 = unpack_results(packed, 5, 7)
    assert torch.equal(c2[:3], count) and int(c2[3]) == 0 and torch.equal(r2[:3], rows)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from ultralytics_pro_b200.dist import shard_range, gather_results
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
B, max_det, cols = 5, 4, 6
g = torch.Generator().manual_seed(0)
full_rows = torch.randn(B, max_det, cols, generator=g)
full_count = torch.tensor([4, 0, 3, 1, 2], dtype=torch.int32)
lo, hi = shard_range(B, rank, world)
rows, count = gather_results(full_rows[lo:hi].clone(), full_count[lo:hi].clone(), B)
assert torch.equal(count, full_count), (rank, count)
assert torch.equal(rows, full_rows), rank
work, finish = gather_results(full_rows[lo:hi].clone(), full_count[lo:hi].clone(), B, async_op=True)
work.wait()
rows, count = finish()
assert torch.equal(count, full_count) and torch.equal(rows, full_rows)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gather_gloo(tmp_path):
    """World size 2 on CPU: contiguous sharding + the packed all_gather reproduce the unsharded results on every rank."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29561", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_patch_points_exist_in_reference_when_mounted():
    """In the build container the reference tree is mounted: install() must find and rebind every documented symbol."""
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    ref = load_reference()
    import ultralytics_pro_b200.patch as patch

    done = patch.install()
    try:
        assert "ultralytics.utils.nms.non_max_suppression" in done
        assert "ultralytics.nn.modules.head.Detect._inference" in done
        assert {"ultralytics.utils.nms.TorchNMS.nms", "ultralytics.utils.nms.TorchNMS.fast_nms"} <= set(done)
        # CPU tensors still reach the untouched reference
        y = torch.zeros(1, 84, 16)
        y[0, :4, 3] = torch.tensor([10.0, 10.0, 4.0, 4.0])
        y[0, 4, 3] = 0.9
        out = ref.nms.non_max_suppression(y, 0.25, 0.7)
        assert out[0].shape == (1, 6)
        keep = ref.TorchNMS.nms(torch.tensor([[0.0, 0, 10, 10], [1, 1, 11, 11]]), torch.tensor([0.9, 0.8]), 0.5)
        assert keep.tolist() == [0]
        assert {f"ultralytics.utils.ops.{n}" for n in ("scale_boxes", "clip_boxes", "scale_coords", "clip_coords", "regularize_rboxes")} <= set(done)
        bx = ref.ops.scale_boxes((640, 640), torch.tensor([[10.0, 20.0, 700.0, 300.0]]), (480, 640))  # CPU -> reference
        assert bx.tolist() == [[10.0, 0.0, 640.0, 220.0]]
        assert ref.ops.regularize_rboxes(torch.tensor([[1.0, 2.0, 3.0, 4.0, 2.0]]))[0, 2].item() == 4.0
    finally:
        patch.uninstall()
    assert ref.nms.non_max_suppression.__name__ == "non_max_suppression" and not hasattr(ref.nms.non_max_suppression, "__wrapped__")


def test_new_entry_points_validate_arguments_before_any_launch():
    """The 8f / multi-GPU entry points reject bad arguments on the host (no device needed) with the documented codes."""
    import ctypes as C

    from ultralytics_pro_b200 import _cabi

    lib = _cabi.load()
    err = lambda: lib.ypb_last_error_string().decode()  # noqa: E731
    xf = _cabi.ScaleXform()
    assert lib.ypb_scale_rows(None, 0, 4, 1, 0, None, None, C.byref(xf), _cabi.BOXES_XYXY, 0, 0, None, 0, 0, 0, 0, None) == 0  # empty: OK
    assert lib.ypb_scale_rows(16, 0, 4, 1, 3, None, None, C.byref(xf), 17, 0, 0, None, 0, 0, 0, 0, None) == -1 and "box mode" in err()
    assert lib.ypb_scale_rows(16, 0, 4, 1, 3, None, None, None, _cabi.BOXES_XYXY, 0, 0, None, 0, 0, 0, 0, None) == -1 and "transform" in err()
    assert lib.ypb_scale_rows(16, 0, 5, 1, 3, None, None, C.byref(xf), _cabi.BOXES_XYWHR, 0, 2, None, 0, 0, 0, 0, None) == -1 and "angle_col" in err()
    g = _cabi.HeadDesc()
    g.num_levels, g.batch, g.dtype = 1, 1, _cabi.YPB_F32
    g.level_h[0], g.level_w[0], g.level_stride[0] = 4, 4, 8.0
    assert lib.ypb_kpts_decode(C.byref(g), None, 0, 16, 7, 3, None, None) == -1 and "ndim" in err()   # 7 channels, ndim 3
    assert lib.ypb_kpts_decode(C.byref(g), None, 0, 16, 6, 3, None, None) == -1 and "NULL" in err()
    pd = _cabi.ProtosDesc()
    pd.dtype, pd.channels, pd.mh, pd.mw, pd.stride_c = _cabi.YPB_F32, 32, 8, 8, 64
    args = [None, 0, 32, None, 0, 4, None, 1, 0, 16, 16, 0, 0, 8, 8, _cabi.MASK_CROP_PROTO, 1.0, 1.0, None, None, 0, None]
    assert lib.ypb_process_mask(C.byref(pd), *args) == 0                                            # no detections: OK
    bad = list(args); bad[13] = 9                                                                  # window taller than the grid
    assert lib.ypb_process_mask(C.byref(pd), *bad) == -1 and "window" in err()
    bad = list(args); bad[15] = 7
    assert lib.ypb_process_mask(C.byref(pd), *bad) == -1 and "crop mode" in err()
    thr = (C.c_float * 2)(0.5, 0.75)
    assert lib.ypb_match_predictions(None, 0, 6, 5, 1, 0, None, None, None, 0, 0, None, 0, None, thr, 2, None, None, 0, None) == 0
    assert lib.ypb_match_predictions(None, 0, 6, 5, 1, 4, None, None, None, 0, 0, None, 0, None, thr, 0, None, None, 0, None) == -1
    assert lib.ypb_peer_wait(None, 2, None, 0, 3, None, 0, None, None) == -1
    # ABI structs the header documents
    assert C.sizeof(_cabi.ScaleXform) == 32 and _cabi.MAX_PEERS == 8
    out = _cabi.NmsOut()
    out.num_peers = 9
    p = _cabi.NmsParams()
    p.nc, p.max_det, p.max_nms, p.rows_cap, p.conf_thres, p.iou_thres_eff = 1, 1, 1, 1, 0.25, 0.5
    out.rows = out.count = 1  # non-NULL dummies: validation stops at num_peers before any use
    d = _cabi.DenseDesc()
    d.dtype, d.batch, d.channels, d.anchors = _cabi.YPB_F32, 1, 5, 4
    assert lib.ypb_nms_from_dense(C.byref(d), C.byref(p), C.byref(out), None, 0, None) == -1 and "num_peers" in err()


def test_c_program_links_against_the_library(tmp_path):
    """A plain C99 translation unit includes the header, links the shared library and exercises the host-only entry points."""
    import shutil
    import subprocess

    from ultralytics_pro_b200 import build

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    lib = build.build_library()
    exe = tmp_path / "c_abi_smoke"
    src = os.path.join(ROOT, "tests", "c_abi_smoke.c")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", str(exe),
           "-L", os.path.dirname(lib), "-lyolopost_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.run(cmd, check=True, capture_output=True)
    env = dict(os.environ)
    cudart = "/usr/local/cuda/lib64"
    env["LD_LIBRARY_PATH"] = cudart + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c abi ok" in out.stdout


def test_host_scalars_of_the_result_ops_match_the_oracle():
    """letterbox gain / pad scalars (ops.py:120-127, :580-587) and the scale_masks window (ops.py:544-559): the host side of
    the drop-ins computes them in Python - they must equal the oracle's (which is pinned to the live reference)."""
    import struct

    from oracle import result_ops_oracle as ro
    from ultralytics_pro_b200 import ops

    f32 = lambda v: struct.unpack("f", struct.pack("f", v))[0]  # noqa: E731
    cases = [((640, 640), (480, 640, 3), None), ((384, 640), (1080, 1920), None), ((640, 480), (1333, 999), None),
             ((1024, 1024), (3000, 4000, 3), None), ((640, 640), (500, 375), ((1.28, 1.28), (80.5, 0.25)))]
    for img1, img0, rp in cases:
        gain, pad, cpad = ro.letterbox_scalars(img1, img0, rp)
        x = ops.letterbox_transform(img1, img0, rp)
        assert (x.gain, x.pad_x, x.pad_y, x.cpad_x, x.cpad_y) == (f32(gain), f32(pad[0]), f32(pad[1]), f32(cpad[0]), f32(cpad[1]))
        assert (x.img_w, x.img_h) == (float(img0[1]), float(img0[0]))
    for mh, mw, shape in [(160, 160, (480, 640)), (40, 40, (120, 213)), (40, 40, (333, 250)), (96, 160, (720, 1280))]:
        top, left, bottom, right = ro.scale_masks_window(mh, mw, shape)
        assert ops._scale_masks_window(mh, mw, shape) == (top, left, bottom - top, right - left)


def test_new_drop_ins_refuse_cpu_tensors():
    from ultralytics_pro_b200 import export_nms, head, ops, val

    z = torch.zeros
    for call in (lambda: ops.clip_boxes(z(2, 4), (10, 10)), lambda: ops.scale_coords((8, 8), z(2, 3), (4, 4)),
                 lambda: ops.regularize_rboxes(z(2, 5)), lambda: ops.process_mask(z(4, 8, 8), z(1, 4), z(1, 4), (32, 32)),
                 lambda: head.decode_keypoints(z(1, 6, 16), [(4, 4)], [8], (2, 3)), lambda: head.detect_postprocess(z(1, 16, 9), 5, 5),
                 lambda: val.match_iou_matrix([0.5], z(2), z(2), z(2, 2)), lambda: export_nms.nms_model_postprocess(z(1, 9, 16), (32, 32), 5, 0.25, 0.5, 5)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
