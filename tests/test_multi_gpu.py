"""One-sided peer-memory gather (dist.PeerGather, ypb_nms_out.peer_*) against an NCCL all_gather of the same results.

Needs >= 2 GPUs on the box (skipped otherwise; the world-size-2 host logic is covered on CPU with gloo in
tests/test_host_logic.py).  Two ranks, one process per GPU, each post-processes its own shard through
HeadPostProcessor(peer_gather_group=True) - eagerly, with lag, and replayed from a CUDA graph."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

pytestmark = pytest.mark.gpu

SCRIPT = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["YPB_ROOT"])
    import torch, torch.distributed as dist
    from ultralytics_pro_b200 import dist as ypb_dist
    from ultralytics_pro_b200.pipeline import HeadPostProcessor
    from ultralytics_pro_b200.synth import HeadConfig, make_head_batch
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", rank); torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = HeadConfig("mg", 160, (8, 16, 32), 80, 4, objects=6)
    B = 4
    pp = HeadPostProcessor(cfg.nc, cfg.strides, 0.25, 0.7, peer_gather_group=True)
    st = torch.cuda.Stream(dev)
    def check(levels, tag):
        pl = pp.last
        torch.cuda.synchronize(dev); dist.barrier()
        ref = torch.empty((world, pl.packed.numel()), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, pl.packed.clone())
        rr, rc = ypb_dist.split_packed(ref, B, pl.rows.shape[1], pl.rows.shape[2])
        gr, gc = pp.gathered()
        valid = (torch.arange(pl.rows.shape[1], device=dev)[None, :] < rc[:, None]).unsqueeze(-1)
        assert torch.equal(rc, gc), (tag, rc.tolist(), gc.tolist())
        assert torch.equal(rr * valid, gr * valid), tag
        assert int(rc.sum()) > 0, tag
    with torch.cuda.stream(st):
        for step in range(3):  # eager, different data each step and rank
            lv = [t.to(dev) for t in make_head_batch(cfg, batch=B, seed=10 * step + rank, first_image=rank * B)[0]]
            pp.enqueue(lv); pp.wait_gather(0)
            check(lv, f"eager {step}")
        # pipelined (lag 1) and CONSUMED every step: step k reads the ring entry of batch k-1 (device-side slot index, no host
        # sync) while the producers already store batch k; 12 steps > 3 ring entries, so entries are reused under the
        # acknowledgement protocol.  Every consumed batch must equal the NCCL all_gather of that batch.
        sets = [[t.to(dev) for t in make_head_batch(cfg, batch=B, seed=100 + 7 * s + rank, first_image=rank * B)[0]] for s in range(4)]
        consumed, wanted = [], []
        for step in range(12):
            pl = pp.enqueue(sets[step % 4])
            ref = torch.empty((world, pl.packed.numel()), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(ref, pl.packed.clone())
            wanted.append(ref)
            pp.wait_gather(1)
            if step >= 1:
                consumed.append(pl.peers.entry_tensor().clone())
        pp.wait_gather(0)
        consumed.append(pp.last.peers.entry_tensor().clone())
        torch.cuda.synchronize(dev)
        nrow, md, cols = pp.last.peers.nrow, pp.last.rows.shape[1], pp.last.rows.shape[2]
        for k, (got, ref) in enumerate(zip(consumed, wanted)):
            rr, rc = ypb_dist.split_packed(ref, B, md, cols)
            gr, gc = ypb_dist.split_packed(got.contiguous(), B, md, cols)
            valid = (torch.arange(md, device=dev)[None, :] < rc[:, None]).unsqueeze(-1)
            assert torch.equal(rc, gc), ("lag-1 consume", k, rc.tolist(), gc.tolist())
            assert torch.equal(rr * valid, gr * valid), ("lag-1 consume", k)
        # the same inside CUDA graphs (kernels + wait + consumer copy captured together)
        lv = sets[0]
        peers = pp.last.peers
        out = torch.zeros((world, peers.slot), dtype=torch.float32, device=dev)
        g = pp.capture(lv, after=lambda: peers.wait_copy(out, 1))  # wait + consumer copy in ONE captured kernel
        for _ in range(7):
            g.replay()
        pp.wait_gather(0)
        check(lv, "graph, lag 1 + drain")
        peers.wait_copy(out, 0)
        torch.cuda.synchronize(dev)
        rr, rc = ypb_dist.split_packed(out[:, : peers.numel].contiguous(), B, md, cols)
        gr, gc = pp.gathered()
        assert torch.equal(rc, gc) and int(gc.sum()) > 0, "wait_copy consumer"
        # the consumer as a FORKED branch of the NEXT step's graph (in-order wait, lag -1): replay r consumes - beside its own
        # kernels - the batch the previous launch produced; 9 launches > 3 ring entries, so entries are reused under acknowledgements
        plain = HeadPostProcessor(cfg.nc, cfg.strides, 0.25, 0.7)
        refs = []
        for s_ in range(4):
            q = plain.enqueue(sets[s_])
            r_ = torch.empty((world, q.packed.numel()), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(r_, q.packed.clone())
            refs.append(r_)
        gb = torch.zeros((1, world, peers.slot), dtype=torch.float32, device=dev)
        def consume():
            pp.wait_gather(-1)
            peers.copy_entry(gb)
        pp.enqueue(sets[3])  # one batch in flight: the steady state of a lag-1 pipeline
        snaps, expect = [], []
        graphs = []
        for i in (0, 1):
            graphs.append(pp.capture(sets[i], beside=consume))  # eager warm-up inside: consumes the batch in flight, launches set i
            snaps.append(gb.clone()); expect.append(3 if i == 0 else 0)
        last = 1
        for r_ in range(7):
            graphs[r_ % 2].replay()
            snaps.append(gb.clone()); expect.append(last)
            last = r_ % 2
        pp.wait_gather(0)
        torch.cuda.synchronize(dev)
        for k, (got, e_) in enumerate(zip(snaps, expect)):
            rr, rc = ypb_dist.split_packed(refs[e_], B, md, cols)
            gr, gc = ypb_dist.split_packed(got[0][:, : peers.numel].contiguous(), B, md, cols)
            valid = (torch.arange(md, device=dev)[None, :] < rc[:, None]).unsqueeze(-1)
            assert torch.equal(rc, gc), ("forked consumer", k, e_, rc.tolist(), gc.tolist())
            assert torch.equal(rr * valid, gr * valid), ("forked consumer", k)
        rr, rc = ypb_dist.split_packed(refs[last], B, md, cols)
        gr, gc = pp.gathered()
        assert torch.equal(rc, gc) and int(gc.sum()) > 0, "forked consumer drain"
        assert peers.overrun() == 0, peers.overrun()
    print("ok", rank, flush=True)
    dist.barrier(); torch.cuda.synchronize(dev); os._exit(0)
''')


def test_peer_gather_matches_nccl_all_gather(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "peer.py"
    script.write_text(SCRIPT)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", WORLD_SIZE="2", YPB_ROOT=root)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    try:
        outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs), outs
