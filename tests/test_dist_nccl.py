"""The multi-GPU path on real GPUs: two NCCL ranks shard a volume (prototype broadcast, sharded match, compact record
gather) and rank 0 checks the gathered records byte for byte against its own single-rank run.  Needs >= 2 GPUs (skipped
otherwise; run with `gpurun --gpus 2`).  The same host logic runs over gloo on CPU in tests/test_dist_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_total, p2p, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from protosam_b200 import ops, synth
        from protosam_b200.engine import CoarseVolumeEngine, GraphedVolumeStep, shard_range
        cfg = synth.CONFIGS["cfg2_chaos_mri"]
        L = 2
        vol = synth.make_volume(99, Q=q_total, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
        sup, fg = torch.from_numpy(vol.sup).to(dev), torch.from_numpy(vol.fg).to(dev)
        lo, hi = shard_range(q_total, world, rank)
        mine = torch.from_numpy(vol.qry[lo:hi]).to(dev)
        eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"], p2p=p2p)
        if rank != 0:
            sup = torch.zeros_like(sup)              # only the source rank's support features may matter
        eng.set_support(sup, fg, src=0)
        hdr_all, recs_all = eng.run_sharded(mine, q_total, dst=0)
        # the CUDA-graph form of the same step, twice in a row without waiting in between (the step itself must order
        # the second replay after the first gather)
        gs = GraphedVolumeStep(eng, sup, fg, mine, q_total=q_total, src=0)
        gs.launch()
        h2, r2 = gs.launch().result()
        torch.cuda.synchronize()
        # source and destination ranks may change from one exchange to the next (bench.py's lanes use src = lane % ranks)
        sup1 = torch.from_numpy(vol.sup).to(dev) if rank == 1 else torch.zeros_like(sup)
        eng.set_support(sup1, fg, src=1)
        h3, r3 = eng.run_sharded(mine, q_total, dst=1)
        eng.set_support(sup, fg, src=0)
        h4, r4 = eng.run_sharded(mine, q_total, dst=0)
        torch.cuda.synchronize()
        swapped_ok = (h3 is not None) == (rank == 1) and (h4 is not None) == (rank == 0)
        if rank == 1:
            swapped_ok &= h3.shape[0] == q_total * L
        if rank == 0:
            swapped_ok &= torch.equal(h4, hdr_all) and torch.equal(r4, recs_all)
        if rank == 0:
            solo = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"])
            solo.set_support(sup, fg, broadcast=False)
            hs, rs = solo.run(torch.from_numpy(vol.qry).to(dev))
            Hs, Rs = ops.decode_headers(hs), ops.decode_records(rs)
            ok = True
            for H, R in ((hdr_all, recs_all), (h2, r2)):
                Hg = ops.decode_headers(H)
                Rg = np.frombuffer(R.cpu().numpy().tobytes(), dtype=ops.REC_DTYPE)
                ok &= len(Hg) == q_total * L and len(Rg) == int(Hs["n_rec"].sum())
                for i in range(q_total * L):
                    k, a = int(Hs["n_rec"][i]), int(Hg["reserved"][i])
                    ok &= all(np.array_equal(Hg[f][i], Hs[f][i]) for f in Hs.dtype.names if f != "reserved")
                    ok &= Rg[a: a + k].tobytes() == Rs[i, :k].tobytes()
            a = solo.decode(hs, rs)
            b = solo.decode(hdr_all, recs_all)
            ok &= all(x.empty == y.empty and (x.empty or np.array_equal(x.points, y.points)) for x, y in zip(sum(a, []), sum(b, [])))
            out_q.put(bool(ok) and bool(swapped_ok))
        else:
            out_q.put(hdr_all is None and recs_all is None and h2 is None and bool(swapped_ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("q_total", [4, 5])
def test_sharded_volume_over_nccl_equals_single_rank(q_total, p2p):
    """p2p: the table and the records move with the library's one-sided peer-memory kernels (psam_peer_*) instead of the
    NCCL broadcast / gather; the gathered records must be the same bytes either way."""
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_total, p2p, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res)
