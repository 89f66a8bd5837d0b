"""Seeded synthetic Detect-head outputs of the BASELINE.json shapes (SURVEY.md section 8d "clustered objects").

Per image: K objects (uniform centres, log-uniform sizes, uniform class).  Anchors whose centre falls in the central
quarter of an object, and whose ltrb distances fit the DFL range, get box logits peaked at the true distances and a
high logit for the object's class; every other logit is background noise.  That yields spatially clustered candidate
sets with heavy mutual overlap - the regime NMS actually works in - instead of iid noise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch


@dataclass(frozen=True)
class HeadConfig:
    name: str
    imgsz: int
    strides: tuple
    nc: int
    batch: int
    reg_max: int = 16
    rotated: bool = False
    conf: float = 0.25
    iou: float = 0.7
    multi_label: bool = False
    agnostic: bool = False
    max_det: int = 300
    max_nms: int = 30000
    mu_bg: float = -7.0
    objects: int = 30

    @property
    def level_hw(self):
        return tuple((self.imgsz // s, self.imgsz // s) for s in self.strides)

    @property
    def anchors(self):
        return sum(h * w for h, w in self.level_hw)

    @property
    def no(self):
        return 4 * self.reg_max + self.nc


# BASELINE.json configs[0..4]
CONFIGS = {
    "c1_v8n_640_b1": HeadConfig("c1_v8n_640_b1", 640, (8, 16, 32), 80, 1),
    "c2_v8x_640_b64": HeadConfig("c2_v8x_640_b64", 640, (8, 16, 32), 80, 64),
    "c3_val_stress_b32": HeadConfig("c3_val_stress_b32", 640, (8, 16, 32), 80, 32, conf=0.001, multi_label=True, mu_bg=-9.5),
    "c4_p6_1280_b16": HeadConfig("c4_p6_1280_b16", 1280, (8, 16, 32, 64), 80, 16),
    "c5_obb_1024_b16": HeadConfig("c5_obb_1024_b16", 1024, (8, 16, 32), 15, 16, rotated=True),
}


def make_head_batch(cfg: HeadConfig, batch: int | None = None, seed: int = 0, device="cpu", dtype=torch.float32,
                    first_image: int = 0):
    """Returns (levels, angle_logits): list of (B, no, H, W) tensors and (B, 1, A) or None.

    Image i of the batch depends only on (seed, first_image + i), so a batch sharded over ranks is the same data as
    the unsharded batch.
    """
    bsz = cfg.batch if batch is None else batch
    dev = torch.device(device)
    R, nc, K = cfg.reg_max, cfg.nc, cfg.objects
    levels = [torch.empty((bsz, cfg.no, h, w), dtype=torch.float32, device=dev) for h, w in cfg.level_hw]
    angle = torch.empty((bsz, 1, cfg.anchors), dtype=torch.float32, device=dev) if cfg.rotated else None
    bins = torch.arange(R, dtype=torch.float32, device=dev).view(1, R, 1)
    for bi in range(bsz):
        g = torch.Generator(device=dev)
        g.manual_seed((seed * 1_000_003 + first_image + bi) & 0x7FFFFFFF)
        u = lambda *s: torch.rand(*s, generator=g, device=dev)
        n = lambda *s: torch.randn(*s, generator=g, device=dev)
        cxy = u(K, 2) * cfg.imgsz
        wh = torch.exp(u(K, 2) * (math.log(0.6 * cfg.imgsz) - math.log(12.0)) + math.log(12.0))
        cls = torch.randint(0, nc, (K,), generator=g, device=dev)
        obj_ang = (u(K) - 0.25) * math.pi if cfg.rotated else None
        a0 = 0
        for li, ((h, w), s) in enumerate(zip(cfg.level_hw, cfg.strides)):
            hw = h * w
            ax = ((torch.arange(w, device=dev, dtype=torch.float32) + 0.5) * s).repeat(h)
            ay = ((torch.arange(h, device=dev, dtype=torch.float32) + 0.5) * s).repeat_interleave(w)
            # ltrb distances (grid units) from each anchor to each object's sides: (K, hw)
            l = (ax[None] - (cxy[:, :1] - wh[:, :1] / 2)) / s
            t = (ay[None] - (cxy[:, 1:] - wh[:, 1:] / 2)) / s
            r = ((cxy[:, :1] + wh[:, :1] / 2) - ax[None]) / s
            b = ((cxy[:, 1:] + wh[:, 1:] / 2) - ay[None]) / s
            inside = ((ax[None] - cxy[:, :1]).abs() < wh[:, :1] / 4) & ((ay[None] - cxy[:, 1:]).abs() < wh[:, 1:] / 4)
            fits = (torch.stack((l, t, r, b)).amax(0) < R - 1) & (torch.stack((l, t, r, b)).amin(0) > 0)
            pos = inside & fits  # (K, hw)
            owner = torch.where(pos.any(0), pos.float().argmax(0), torch.full((hw,), -1, device=dev))
            is_pos = owner >= 0
            oi = owner.clamp(min=0)
            box = n(4 * R, hw)
            clsl = n(nc, hw) * 1.5 + cfg.mu_bg
            if is_pos.any():
                d = torch.stack((l, t, r, b))[:, oi, torch.arange(hw, device=dev)]  # (4, hw)
                d = d + n(4, hw) * 0.3
                peak = (-2.0 * (bins - d.view(4, 1, hw)) ** 2).view(4 * R, hw)
                box = torch.where(is_pos[None], peak, box)
                obj_logit = n(hw) * 1.5 + 1.0
                sel = torch.zeros((nc, hw), dtype=torch.bool, device=dev)
                sel[cls[oi], torch.arange(hw, device=dev)] = True
                clsl = torch.where(sel & is_pos[None], obj_logit[None], clsl)
            levels[li][bi, : 4 * R] = box.view(4 * R, h, w)
            levels[li][bi, 4 * R:] = clsl.view(nc, h, w)
            if cfg.rotated:
                al = n(hw)
                if is_pos.any():
                    target = obj_ang[oi] + n(hw) * 0.05
                    p = (target / math.pi + 0.25).clamp(1e-3, 1 - 1e-3)
                    al = torch.where(is_pos, torch.log(p / (1 - p)), al)
                angle[bi, 0, a0:a0 + hw] = al
            a0 += hw
    if dtype != torch.float32:
        levels = [lv.to(dtype) for lv in levels]
        angle = angle.to(dtype) if angle is not None else None
    return levels, angle


def make_uniform_batch(cfg: HeadConfig, batch: int | None = None, seed: int = 0, device="cpu", dtype=torch.float32,
                       sigma: float = 2.0):
    """iid N(0, sigma^2) logits: decouples the HBM measurement of the decode from data-dependent NMS work."""
    bsz = cfg.batch if batch is None else batch
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    levels = [(torch.randn((bsz, cfg.no, h, w), generator=g, device=device) * sigma).to(dtype) for h, w in cfg.level_hw]
    angle = torch.randn((bsz, 1, cfg.anchors), generator=g, device=device).to(dtype) if cfg.rotated else None
    return levels, angle
