"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the
prototype broadcast, the slice sharding and the prompt-record gather of protosam_b200.engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from protosam_b200 import engine, ops, prompts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_total, L, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) broadcast of the packed prototype tables (rank 0 holds the real ones)
        nsets, cap, C = 2 * L, 9, 16
        g = torch.Generator().manual_seed(5)
        truth = dict(protos=torch.randn(nsets, cap, C, generator=g),
                     counts=torch.arange(nsets, dtype=torch.int32), eff_modes=torch.ones(nsets, dtype=torch.int32),
                     status=torch.zeros(nsets, dtype=torch.int32))
        mine = {k: (v.clone() if rank == 0 else torch.zeros_like(v)) for k, v in truth.items()}
        engine.broadcast_prototypes(mine, src=0)
        ok_bcast = all(torch.equal(mine[k], truth[k]) for k in truth)
        # 2) shard + gather of fixed-size records, ragged split (q_total not divisible by world)
        lo, hi = engine.shard_range(q_total, world, rank)
        n_local = (hi - lo) * L
        max_cc = 3
        hdr = np.zeros(n_local, ops.HDR_DTYPE)
        recs = np.zeros((n_local, max_cc), ops.REC_DTYPE)
        for i in range(n_local):
            gi = lo * L + i
            hdr["ncc"][i] = hdr["n_rec"][i] = 1 + gi % max_cc
            recs["label"][i, : 1 + gi % max_cc] = np.arange(1, 2 + gi % max_cc)
            recs["box"][i, :, 0] = gi
            recs["centroid"][i, :, 0] = gi + 0.5
        th = torch.from_numpy(hdr.view(np.uint8).reshape(n_local, 64).copy())
        tr = torch.from_numpy(recs.view(np.uint8).reshape(n_local, max_cc, 96).copy())
        counts = [(b - a) * L for a, b in (engine.shard_range(q_total, world, r) for r in range(world))]
        H, R = engine.gather_records(th, tr, counts, dst=0)
        # 3) the engine's own layout: ONE packed table broadcast, ONE packed asynchronous gather
        tab = ops.proto_table_alloc(nsets, cap, C, "cpu")
        if rank == 0:
            for k in ("protos", "counts", "eff_modes", "status"):
                tab[k].copy_(truth[k])
        else:
            tab["packed"].zero_()
        engine.broadcast_prototypes(tab, src=0)
        ok_bcast &= all(torch.equal(tab[k], truth[k]) for k in truth)
        buf, bh, br = ops.records_alloc(max(counts), max_cc, "cpu")
        bh[:n_local] = th
        br[:n_local] = tr
        pend = engine.gather_packed(buf, counts, max_cc, dst=0, async_op=True)
        H2, R2 = pend.result()
        if rank == 0:
            ok_bcast &= torch.equal(H2, H) and torch.equal(R2, R)
        else:
            ok_bcast &= H2 is None and R2 is None
        # 4) the compact layout the engine gathers: [n_alloc headers | tail | capacity live records] per rank
        n_alloc, cap = max(counts), max(counts) * max_cc
        cbuf = torch.zeros(ops.packed_bytes(n_alloc, cap), dtype=torch.uint8)
        ch, ct, cr = ops.split_packed(cbuf, n_alloc, cap)
        first = np.concatenate([[0], np.cumsum(hdr["n_rec"])]).astype(np.int32)
        hc = hdr.copy()
        hc["reserved"] = first[:-1]
        ch[:n_local] = torch.from_numpy(hc.view(np.uint8).reshape(n_local, 64).copy())
        live = np.concatenate([recs[i, : hdr["n_rec"][i]] for i in range(n_local)]) if n_local else np.zeros(0, ops.REC_DTYPE)
        cr[: len(live)] = torch.from_numpy(live.view(np.uint8).reshape(len(live), 96).copy())
        tail = np.zeros(1, ops.TAIL_DTYPE)
        tail["total"], tail["capacity"], tail["n_img"] = len(live), cap, n_local
        ct.copy_(torch.from_numpy(tail.view(np.uint8).reshape(64).copy()))
        H3, R3 = engine.gather_packed(cbuf, counts, ("compact", n_alloc, cap), dst=0, async_op=True).result()
        if rank == 0:
            Hc = ops.decode_headers(H3)
            Rc = np.frombuffer(R3.numpy().tobytes(), dtype=ops.REC_DTYPE)
            okc = len(Hc) == q_total * L and len(Rc) == int(Hc["n_rec"].sum())
            for gi in range(q_total * L):
                a, k = int(Hc["reserved"][gi]), int(Hc["n_rec"][gi])
                okc &= k == 1 + gi % max_cc and bool((Rc["box"][a: a + k, 0] == gi).all())
                okc &= bool(np.array_equal(Rc["label"][a: a + k], np.arange(1, k + 1)))
            ok_bcast &= bool(okc)
        else:
            ok_bcast &= H3 is None and R3 is None
        if rank == 0:
            Hn, Rn = ops.decode_headers(H), ops.decode_records(R)
            ok = len(Hn) == q_total * L
            for gi in range(q_total * L):
                ok &= int(Hn["n_rec"][gi]) == 1 + gi % max_cc and int(Rn["box"][gi, 0, 0]) == gi
                ok &= float(Rn["centroid"][gi, 0, 0]) == gi + 0.5
                sp = prompts.prompts_from_records(Hn[gi], Rn[gi], use_cca=False, point_mode="both")
                ok &= len(sp.predict_calls()) == 1 + gi % max_cc
            out_q.put((ok_bcast, bool(ok)))
        else:
            assert H is None and R is None
            out_q.put((ok_bcast, True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("q_total", [4, 5])
def test_broadcast_shard_gather_world2(q_total):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_total, 2, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in res)


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 1024):
        for world in (1, 2, 3, 8):
            blocks = [engine.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_part_assign_matches_reference_rule():
    """dataloaders/common.py:241-249: int((z - z_min) // ((z_max - z_min) / npart)), clamped, 0 on a degenerate extent"""
    for z_min, z_max, npart in ((10, 40, 3), (0, 7, 3), (5, 5, 3), (3, 50, 4), (12, 13, 3)):
        for z in range(z_min - 3, z_max + 4):
            try:
                ref = int((z - z_min) // ((z_max - z_min) / npart))
            except ZeroDivisionError:
                ref = 0
            ref = 0 if ref < 0 else (npart - 1 if ref >= npart else ref)
            assert engine.part_assign(z, z_min, z_max, npart) == ref


def test_packed_layouts_round_trip_on_cpu():
    """the one-buffer layouts the collectives move: views alias the packed storage, split inverts alloc"""
    tab = ops.proto_table_alloc(4, 7, 16, "cpu")
    tab["protos"].fill_(1.5); tab["counts"].copy_(torch.arange(4, dtype=torch.int32)); tab["status"].zero_(); tab["eff_modes"].fill_(2)
    clone = tab["packed"].clone()
    t2 = ops.proto_table_alloc(4, 7, 16, "cpu")
    t2["packed"].copy_(clone)
    assert torch.equal(t2["protos"], tab["protos"]) and torch.equal(t2["counts"], tab["counts"])
    assert torch.equal(t2["eff_modes"], tab["eff_modes"]) and tab["protos"].shape == (4, 7, 16)
    buf, hdr, recs = ops.records_alloc(5, 3, "cpu")
    hdr[2, 0] = 7
    recs[4, 2, 95] = 9
    h2, r2 = ops.split_records(buf.clone(), 3)
    assert h2.shape == (5, 64) and r2.shape == (5, 3, 96) and h2[2, 0] == 7 and r2[4, 2, 95] == 9
    p = engine.PendingGather.done((hdr, recs))
    assert p.result()[0] is hdr
