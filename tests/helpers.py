"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import torch

from ultralytics_pro_b200.synth import HeadConfig, make_head_batch
from oracle.postproc_oracle import decode_oracle, obb_forward_oracle


def small_cfg(name="small", imgsz=160, nc=80, batch=2, **kw) -> HeadConfig:
    return HeadConfig(name, imgsz, (8, 16, 32), nc, batch, **kw)


def dense_from_oracle(cfg: HeadConfig, batch: int, seed: int, dtype=torch.float32):
    levels, ang = make_head_batch(cfg, batch=batch, seed=seed, dtype=dtype)
    if cfg.rotated:
        y = obb_forward_oracle(levels, ang, cfg.strides, cfg.nc, cfg.reg_max)
    else:
        y = decode_oracle(levels, cfg.strides, cfg.nc, cfg.reg_max)
    return levels, ang, y


def make_scores_unique(y: torch.Tensor, nc: int, floor: float) -> torch.Tensor:
    """Nudge duplicated class scores above `floor` (per image) so ranks are tie-free (Ties)."""
    y = y.clone()
    for b in range(y.shape[0]):
        s = y[b, 4:4 + nc]
        flat = s.reshape(-1)
        idx = torch.nonzero(flat > floor).squeeze(1)
        for _ in range(8):
            vals = flat[idx].float()
            order = torch.argsort(vals, stable=True)
            sv = vals[order]
            dup = torch.zeros_like(sv, dtype=torch.bool)
            dup[1:] = sv[1:] == sv[:-1]
            if not dup.any():
                break
            bump = torch.cumsum(dup.float(), 0) * dup
            nv = sv.clone()
            for _k in range(int(bump.max().item())):
                m = bump > _k
                nv[m] = torch.nextafter(nv[m], torch.full_like(nv[m], 2.0))
            flat[idx[order]] = nv.to(flat.dtype)
        y[b, 4:4 + nc] = flat.view_as(s)
    return y


def assert_rows_equal(got_rows, got_idx, want_rows, want_idx, what=""):
    assert len(got_rows) == len(want_rows), what
    for b, (g, w) in enumerate(zip(got_rows, want_rows)):
        g = g.detach().cpu()
        assert g.shape == w.shape, f"{what} image {b}: kept {g.shape[0]} vs oracle {w.shape[0]}"
        if want_idx is not None and got_idx is not None:
            gi, wi = got_idx[b].detach().cpu().view(-1), want_idx[b].view(-1)
            assert torch.equal(gi, wi), f"{what} image {b}: kept anchor indices differ at {torch.nonzero(gi != wi)[:5].view(-1).tolist()}"
        same = (g == w) | (torch.isnan(g) & torch.isnan(w))
        assert bool(same.all()), f"{what} image {b}: {int((~same).sum())} row values differ, first {torch.nonzero(~same)[:5].tolist()}"
