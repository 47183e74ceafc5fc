#!/usr/bin/env python
"""CTA-level timeline of the volume pipeline (psam_trace_install): which kernel's CTAs sit on which SM when.

    make -C protosam_b200/csrc clean && make -C protosam_b200/csrc TRACE=1 -j4        (trace hooks are compiled out by default)
    python tools/trace_timeline.py run  [--steps 12] [--lanes 4] -> gpurun_out/trace.npy   (needs a GPU)
    python tools/trace_timeline.py show gpurun_out/trace.npy                                  (anywhere)

`show` prints, per kernel, launches / mean duration / mean number of SMs busy, and for the steady-state window how much of
the GEMM's time other kernels' CTAs were resident on the same SM (the overlap the pipeline is after)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REC = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("smid", "<u4"), ("kernel", "<u4"), ("cta", "<u4"), ("pad", "<u4")])
NAMES = {1: "k_match_tc", 2: "k_pack_query(1/16)", 3: "k_blocks_warp", 4: "k_components", 5: "k_classify_blocks",
         6: "k_proto_stage1(1/8)", 7: "k_proto_stage2(1/16)", 8: "k_pack_protos", 9: "k_compact_records"}


def run(steps, lanes, split=0):
    import argparse
    import ctypes

    import torch

    import bench
    from protosam_b200 import _lib, synth
    args = argparse.Namespace(lanes=lanes, algo=0, split_streams=split, graph_collectives=0)
    cfg = dict(synth.CONFIGS["cfg2_chaos_mri"]); cfg["name"] = "cfg2_chaos_mri"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    pipe = bench.Pipeline(cfg, cfg["Q"], cfg["Q"], [cfg["Q"] * cfg["L"]], args, 1, 0, dev, [None] * lanes, bench.N_ROTATE, True)
    for i in range(8):
        pipe.step(i)
    pipe.sync_all()
    cap = 1 << 20
    buf = torch.zeros(cap * REC.itemsize, dtype=torch.uint8, device=dev)
    L = _lib.load()
    assert L.psam_trace_install(ctypes.c_void_p(buf.data_ptr()), buf.numel()) == 0
    for i in range(steps):
        pipe.step(i)
    pipe.sync_all()
    L.psam_trace_install(None, 0)
    a = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=REC)
    n = int(a[0]["t0"])
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/trace.npy", a[1: n + 1])
    print("records", n)


def show(path):
    a = np.load(path)
    a = a[a["t1"] > 0]
    t_min = a["t0"].min()
    t0 = (a["t0"] - t_min).astype(np.float64) * 1e-3
    t1 = (a["t1"] - t_min).astype(np.float64) * 1e-3
    print(f"{len(a)} CTA records over {t1.max():.0f} us, SMs seen: {len(np.unique(a['smid']))}")
    for k in sorted(np.unique(a["kernel"])):
        m = a["kernel"] == k
        print(f"  {NAMES.get(int(k), k):24s} CTAs {m.sum():6d}  mean CTA life {np.mean(t1[m] - t0[m]):8.1f} us  "
              f"total CTA-time {np.sum(t1[m] - t0[m]) / 1e3:8.2f} ms")
    # per-SM residency: fraction of GEMM CTA lifetime during which a B / comp CTA lived on the same SM
    g = a["kernel"] == 1
    span = (t0[g].min(), t1[g].max())
    busy = np.sum(t1[g] - t0[g])
    nsm = len(np.unique(a["smid"]))
    print(f"GEMM CTAs resident {busy / ((span[1] - span[0]) * nsm) * 100:.1f} % of SM-time in [{span[0]:.0f}, {span[1]:.0f}] us")
    for k in (3, 4, 2, 5):
        o = a["kernel"] == k
        ov = 0.0
        for sm in np.unique(a["smid"]):
            gs = np.flatnonzero(g & (a["smid"] == sm))
            os_ = np.flatnonzero(o & (a["smid"] == sm))
            if len(gs) == 0 or len(os_) == 0:
                continue
            # union of the other kernel's intervals on this SM, intersected with the GEMM intervals
            iv = sorted(zip(t0[os_], t1[os_]))
            merged = []
            for s, e in iv:
                if merged and s <= merged[-1][1]:
                    merged[-1][1] = max(merged[-1][1], e)
                else:
                    merged.append([s, e])
            for gi in gs:
                for s, e in merged:
                    ov += max(0.0, min(e, t1[gi]) - max(s, t0[gi]))
        print(f"  {NAMES[k]:24s} resident beside a GEMM CTA for {ov / busy * 100:5.1f} % of the GEMM CTAs' lifetime")
    # launches of the GEMM in time order: start spread and duration
    gs = np.flatnonzero(g)
    order = gs[np.argsort(t0[gs])]
    starts = t0[order]
    cuts = np.flatnonzero(np.diff(starts) > 30.0)
    bounds = np.concatenate([[0], cuts + 1, [len(order)]])
    print("GEMM launches (start us, first->last CTA start spread us, duration us, CTAs):")
    for i in range(len(bounds) - 1):
        idx = order[bounds[i]: bounds[i + 1]]
        print(f"   {t0[idx].min():9.1f}  spread {t0[idx].max() - t0[idx].min():7.1f}  dur {t1[idx].max() - t0[idx].min():7.1f}  n {len(idx)}")
    for k in (3, 4):
        ks = np.flatnonzero(a["kernel"] == k)
        order = ks[np.argsort(t0[ks])]
        cuts = np.flatnonzero(np.diff(t0[order]) > 30.0)
        bounds = np.concatenate([[0], cuts + 1, [len(order)]])
        print(f"{NAMES[k]} launches (start us, spread, duration, CTAs):")
        for i in range(len(bounds) - 1):
            idx = order[bounds[i]: bounds[i + 1]]
            print(f"   {t0[idx].min():9.1f}  spread {t0[idx].max() - t0[idx].min():7.1f}  dur {t1[idx].max() - t0[idx].min():7.1f}  n {len(idx)}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        import argparse
        ap = argparse.ArgumentParser()
        ap.add_argument("cmd")
        ap.add_argument("--steps", type=int, default=12)
        ap.add_argument("--lanes", type=int, default=4)
        ap.add_argument("--split", type=int, default=0)
        ns = ap.parse_args()
        run(ns.steps, ns.lanes, ns.split)
    else:
        show(sys.argv[2])
