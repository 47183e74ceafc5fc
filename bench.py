#!/usr/bin/env python
"""bench.py -- query slices/sec through the ALP match + prompt path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (the C port on the GPU box)

Headline workload (the JSON line's `value`): one step = one synthetic volume of the CHAOS-MRI-shaped config
(BASELINE.json configs[1]) per GPU: kernel 1 over the support slice for 4 labels (8 prototype sets), kernel 2 over 32
query slices, kernel 3 over the 128 resulting coarse maps -> prompt records.  With N ranks every rank takes 32 further
slices of the same volume (weak scaling): one rank computes the prototypes and broadcasts them (NCCL), every rank matches
its slices, the compacted prompt records are gathered to rank 0.  `--lanes` volumes are in flight per GPU on separate
CUDA streams, each replayed from CUDA graphs (engine.GraphedVolumeStep).

`value` is timed with inputs resident in HBM; `e2e` goes through the same engine from pinned host buffers with the
host->device copies and the device->host read of the records inside the timed region.  The line also carries the
roofline of the dominant kernel (timed live with CUDA events around its launch), a CPU baseline (the oracle port, one
core, bounded sample), at N > 1 a parity check of the gathered records against a single-rank run, and -- under
`north_star_runs` -- the two sharded runs BASELINE.json names (strong scaling: the total is fixed, ranks split it):
config 3 (128 slices x 4 labels) and config 5 (the 1024-slice stress sweep).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "cfg2_chaos_mri"
METRIC = "query_slices_per_sec_alp_match_plus_prompts"
UNIT = "slices/s"
DTYPE = "f32 (contraction: bf16x3 split operands, fp32 accumulate in TMEM; everything else fp32 / exact integer)"
N_ROTATE = 4            # distinct input volumes cycled through the timed region (> L2 in total)


def workload_desc(cfg, n_gpus, scaling="weak", q_total=None, p2p=False):
    per_gpu = cfg["Q"] if scaling == "weak" else None
    return {
        "workload": (f"{cfg['name']}: 1 support + {cfg['Q']} query slices {'per GPU' if scaling == 'weak' else 'in total'}, "
                     f"/14 patch features {cfg['h']}x{cfg['w']}x{cfg['C']}, {cfg['L']} labels (bg 'gridconv' + fg "
                     f"'gridconv+'/'mask' decided on device), ws={cfg['ws']}, upsample {cfg['img_size']}->1024, prompts for "
                     f"every component (use_cca=False, point_mode=both)"),
        "slices_per_step_per_gpu": per_gpu, "slices_per_step_total": q_total,
        "labels": cfg["L"],
        "l2_policy": f"{N_ROTATE} distinct query volumes rotate through the timed region "
                     f"({N_ROTATE * cfg['Q'] * cfg['h'] * cfg['w'] * cfg['C'] * 4 / 1e6:.0f} MB of inputs + "
                     f"{cfg['Q'] * cfg['L'] * 0.5:.0f} MB of per-step intermediates > 126 MB L2)",
        "parallelism": f"slices sharded over {n_gpus} GPU(s); prototype table from the source rank (= lane % ranks) to every "
                       "rank + the compacted records to rank 0, " +
                       ("one-sided stores over NVLink peer memory (psam_peer_*)" if p2p else "NCCL broadcast + gather"),
        "lanes": f"{cfg.get('lanes', 1)} volume(s) in flight per GPU on separate CUDA streams",
    }


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="psam_clocks_", suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "50"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """number of lines written so far (to separate warm-up from timed samples)"""
        try:
            self.fh.flush()
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self, first_line=0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path).read().strip().splitlines()[first_line:]:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# CPU legs: the oracle port (C restatement) and, where the reference tree is mounted, the reference itself
# ----------------------------------------------------------------------------------------------
def cpu_slice_label(O, vol, cfg, q, l, fg_mode):
    """One (query slice, label) unit of the reference path on the CPU: two ALP calls
    (grid_proto_fewshot.py:239-259), then ProtoSAM.forward's prompt extraction (ProtoSAM.py:592-635)."""
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None]
    qry = np.transpose(vol.qry[q], (2, 0, 1))[None]
    ks = [cfg["h"] // 8, cfg["w"] // 8]
    bg, _, _, _ = O.alp_forward(qry, sup_x, vol.bg[l][None, :, None], "gridconv", 0.95, ks, isval=True, val_wsize=cfg["ws"])
    fg, _, _, _ = O.alp_forward(qry, sup_x, vol.fg[l][None, :, None], fg_mode[l], 0.95, ks, isval=True, val_wsize=cfg["ws"])
    low = np.concatenate([bg, fg], 1)
    return O.coarse_to_prompts(low, cfg["img_size"], 1024, use_cca=False, point_mode="both")


def cpu_fg_modes(O, vol, cfg):
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))
    ks = (cfg["h"] // 8, cfg["w"] // 8)
    return ["gridconv+" if (O.get_prototypes(sup_x, vol.fg[l][:, None], "gridconv+", ks, 0.95)["pooled"] >= 0.95).any()
            else "mask" for l in range(cfg["L"])]


def run_cpu_port(cfg, n_slices, threads, steps, warmup):
    """-> (slices/s, seconds per step) for the oracle port on `threads` host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from protosam_b200 import synth
    O.lib()
    vol = synth.make_volume(1234, Q=n_slices, L=cfg["L"], C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    modes = cpu_fg_modes(O, vol, cfg)
    units = [(q, l) for q in range(n_slices) for l in range(cfg["L"])]

    def step():
        if threads == 1:
            for q, l in units:
                cpu_slice_label(O, vol, cfg, q, l, modes)
        else:
            with ThreadPoolExecutor(threads) as ex:       # ctypes releases the GIL inside the C oracle
                list(ex.map(lambda u: cpu_slice_label(O, vol, cfg, u[0], u[1], modes), units))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_slices / dt, dt


def run_cpu_reference(cfg, n_slices, threads, steps, warmup):
    """The UNMODIFIED reference (mounted tree, oracle/ref_shims.py) on the same units: per (slice, label) the two
    MultiProtoAsConv calls + cat + F.interpolate of grid_proto_fewshot.py:239-273, then ProtoSAM.forward with a stub coarse
    model returning those logits and the capturing predictor (ProtoSAM.py:559-635).  -> (slices/s, s per step)."""
    import contextlib
    import io

    import torch
    import torch.nn.functional as F

    from oracle import oracle as O
    from oracle import ref_shims
    from protosam_b200 import synth
    torch.set_num_threads(threads)
    alp = ref_shims.load_alpmodule()
    _, PS, _ = ref_shims.load_pipeline()
    vol = synth.make_volume(1234, Q=n_slices, L=cfg["L"], C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    modes = cpu_fg_modes(O, vol, cfg)
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())   # noqa: E731
    with quiet():
        unit = alp.MultiProtoAsConv(proto_grid=[8, 8], feature_hw=[cfg["h"], cfg["w"]])
    sup_x = torch.from_numpy(vol.sup).permute(0, 3, 1, 2)[None, :, None]
    img = torch.from_numpy(synth.uniform(77, (1, 3, cfg["img_size"], cfg["img_size"])))

    def step():
        for q in range(n_slices):
            qry = torch.from_numpy(vol.qry[q]).permute(2, 0, 1)[None, None]
            for l in range(cfg["L"]):
                with ref_shims.cpu_cuda_identity(), quiet(), torch.no_grad():
                    bg = unit(qry, sup_x, torch.from_numpy(vol.bg[l])[None, :, None], "gridconv", 0.95, isval=True,
                              val_wsize=cfg["ws"])[0]
                    fg = unit(qry, sup_x, torch.from_numpy(vol.fg[l])[None, :, None], modes[l], 0.95, isval=True,
                              val_wsize=cfg["ws"])[0]
                    logits = F.interpolate(torch.cat([bg, fg], 1), size=(cfg["img_size"],) * 2, mode="bilinear")
                    model = PS.ProtoSAM(image_size=(1024, 1024), coarse_segmentation_model=ref_shims.FixedLogitsCoarseModel(logits),
                                        num_points_for_sam=1, use_points=True, use_bbox=True, use_cca=False, point_mode="both")
                    model.eval()
                    model(img, ref_shims._NullInput(), degrees_rotate=0)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_slices / dt, dt


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    warm = max(0, min(args.warmup, 1))
    kind, extra = "port", {}
    have_ref = False
    try:
        from oracle import ref_shims
        have_ref = ref_shims.reference_available() and not args.port_only
    except Exception:
        have_ref = False
    if have_ref:
        # the reference itself (Python, single process: torch intra-op threads are its only parallelism), 1 thread -- its own
        # setting, validation_protosam.py:299 -- and all cores; the port beside it for calibration
        n_slices = 2
        v1, d1 = run_cpu_reference(cfg, n_slices, 1, 1, 1)
        vall, dall = run_cpu_reference(cfg, n_slices, cores, 1, 0)
        vp, dp = run_cpu_port(cfg, n_slices, 1, 1, 0)
        kind, val, dt = "reference", max(v1, vall), min(d1, dall)
        extra = {"reference_1_thread": v1, "reference_all_cores": vall, "port_1_thread": vp,
                 "numpy_caveat": "np.array(tensor) at ProtoSAM.py:602 takes ~0.18 s per call under numpy >= 2 (slow __array__ "
                                 "path); the reference pins numpy 1.23.5"}
        sample = f"{n_slices} query slices x {cfg['L']} labels per step, the unmodified reference through oracle/ref_shims.py"
        steps, warm = 1, 1
    else:
        n_slices = max(1, min(cfg["Q"], cores))
        # keep the whole run within a few minutes: one slice (4 labels) costs ~2 s of one core
        val, dt = run_cpu_port(cfg, n_slices, cores, steps, warm)
        sample = (f"{n_slices} query slices x {cfg['L']} labels per step on {cores} threads (C port of the reference path: the "
                  "reference is Python and is not on this box; measured in the build container the port is ~1.9x FASTER than "
                  "the reference itself on one thread, see BASELINE.md section 5)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_desc(cfg, args.gpus),
        "cpu_baseline": dict({"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, **extra),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Pipeline:
    """`lanes` volumes in flight per GPU, one CUDA-graph replay (or eager enqueue) per volume."""

    def __init__(self, cfg, q_local, q_total, counts_all, args, world, rank, dev, groups, n_rotate, use_graphs):
        import torch

        from protosam_b200 import synth
        from protosam_b200.engine import CoarseVolumeEngine, GraphedVolumeStep
        self.torch, self.cfg, self.world, self.rank, self.dev = torch, cfg, world, rank, dev
        self.q_local, self.q_total, self.counts_all, self.groups = q_local, q_total, counts_all, groups
        L, C, h, w = cfg["L"], cfg["C"], cfg["h"], cfg["w"]
        # synthetic inputs: one support slice; n_rotate query volumes per rank.  Volume 0 is the same on every rank,
        # volumes k > 0 differ per rank (seed 1000 k + rank) -- rank 0 can rebuild any rank's volume for the parity check.
        nb = min(q_local, 32)
        vol = synth.make_volume(1234, Q=nb, L=L, C=C, h=h, w=w, img_size=cfg["img_size"])
        self.vol = vol
        self.sup = torch.from_numpy(vol.sup).to(dev)
        self.fg = torch.from_numpy(vol.fg).to(dev)
        base = torch.from_numpy(vol.qry).to(dev)
        if q_local > nb:            # large shards: tile the 32 synthetic slices, every copy with its own noise
            reps = (q_local + nb - 1) // nb
            g0 = torch.Generator(device=dev).manual_seed(77)
            base = (base.repeat(reps, 1, 1, 1)[:q_local] +
                    0.05 * torch.randn((q_local, h, w, C), generator=g0, device=dev)).contiguous()
        self.base = base
        self.qvols = [base] + [self.rank_volume(k, rank) for k in range(1, n_rotate)]
        self.NL = NL = max(1, args.lanes)
        self.engs = [CoarseVolumeEngine((h, w), cfg["img_size"], out_size=1024, val_wsize=cfg["ws"], use_cca=False,
                                        point_mode="both", match_algo=args.algo, group=groups[ln],
                                        p2p=bool(getattr(args, "p2p", 0)) and world > 1) for ln in range(NL)]
        self.split = bool(getattr(args, "split_streams", False)) and use_graphs
        # split: the lane's own stream carries the match stage, the prompt stage runs on a second, higher-priority stream
        self.lanes = [torch.cuda.Stream(device=dev) for _ in range(NL)]
        self.pending = [None] * NL
        self.replayed = 0
        self.graphs = None
        if use_graphs:
            self.graphs = []
            for ln in range(NL):
                with torch.cuda.stream(self.lanes[ln]):
                    mine = [self.qvols[j] for j in range(n_rotate) if j % NL == ln] or [self.qvols[ln % n_rotate]]
                    self.graphs.append([GraphedVolumeStep(self.engs[ln], self.sup, self.fg, qv, q_total=q_total, src=ln % world,
                                                          split_streams=self.split,
                                                          capture_collectives=bool(getattr(args, "graph_collectives", 0)))
                                        for qv in mine])
            torch.cuda.synchronize()

    def rank_volume(self, k, r):
        """query volume k >= 1 of rank r (a pure function of (k, r): every rank can rebuild any rank's inputs)"""
        torch = self.torch
        g = torch.Generator(device=self.dev).manual_seed(1000 * k + r)
        return (self.base.roll(k, 0) + 0.05 * torch.randn(self.base.shape, generator=g, device=self.dev)).contiguous()

    def step(self, i, ev=None, lane=None, eager=False, vol_index=None):
        """One volume: prototypes (+ broadcast), match, prompts, record gather, all on the stream of lane
        i % NL: consecutive volumes are in flight on different streams, so the small launches, the collectives
        and the tail of one volume overlap the big kernels of the next.  The gather of a volume is asynchronous
        (NCCL's stream) and is waited for when its lane is used again / at the end of the region."""
        from protosam_b200.engine import gather_packed
        torch, NL, world = self.torch, self.NL, self.world
        ln = i % NL if lane is None else lane
        e = self.engs[ln]
        with torch.cuda.stream(self.lanes[ln]):
            if self.pending[ln] is not None:
                self.pending[ln].wait()
            if self.graphs is not None and not eager:
                gs = self.graphs[ln][(i // NL) % len(self.graphs[ln])]
                if ev is not None:
                    ev[0].record()
                out = gs.launch()
                self.replayed += gs.n_kernels
                if ev is not None:
                    ev[2].record()
                if world > 1:
                    self.pending[ln] = out
                return out
            e.set_support(self.sup, self.fg, src=ln % world)
            qv = self.qvols[(i if vol_index is None else vol_index) % len(self.qvols)]
            if ev is not None:
                ev[0].record()
            logits = e.match(qv)
            if ev is not None:
                ev[1].record()
            if world == 1:
                out = e.prompts_from_logits(logits)
                if ev is not None:
                    ev[2].record()
                return out
            n_alloc = max(self.counts_all)
            _, _, buf = e.prompts_from_logits(logits, n_alloc=n_alloc, return_packed=True)
            if ev is not None:
                ev[2].record()
            self.pending[ln] = e.collect_records(buf, self.counts_all, ("compact", n_alloc, n_alloc * e.recs_per_image), dst=0,
                                                 async_op=True)
            return self.pending[ln]

    def drain(self):
        """all lanes: wait for the pending gathers, then make the current stream wait for the lanes"""
        torch = self.torch
        for ln in range(self.NL):
            with torch.cuda.stream(self.lanes[ln]):
                if self.pending[ln] is not None:
                    self.pending[ln].wait()
                    self.pending[ln] = None
                for gs in (self.graphs[ln] if self.graphs is not None else []):
                    gs.join()
            torch.cuda.current_stream().wait_stream(self.lanes[ln])

    def sync_all(self):
        import torch.distributed as dist
        self.drain()
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, K, warmup):
        """-> (ms per step: max over ranks, host ms to enqueue a step, per-step events)"""
        import torch.distributed as dist
        torch = self.torch
        for i in range(warmup):
            self.step(i)
        self.sync_all()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        t_start.record()
        for ln in range(self.NL):
            self.lanes[ln].wait_event(t_start)
        th0 = time.perf_counter()
        for i in range(K):
            self.step(i, evs[i])
        host_ms = (time.perf_counter() - th0) * 1e3 / K     # host time to enqueue one step (launch-bound if ~ ms_per_step)
        self.drain()                        # the last gathers are part of the timed region
        t_end.record()
        self.sync_all()
        t = torch.tensor([t_start.elapsed_time(t_end)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / K, host_ms, evs


def shard_counts(q_total, world, L):
    from protosam_b200.engine import shard_range
    return [(b - a) * L for a, b in (shard_range(q_total, world, r) for r in range(world))]


_HDR_SPLIT = np.dtype([("ncc", "<i4"), ("n_rec", "<i4"), ("body", "u1", 52), ("first", "<i4")])   # 64 B: hdr.reserved last


def parity_check(pipe, args):
    """N > 1: the records rank 0 receives through the NCCL path (sharded match, prototype broadcast, compact gather) must
    equal, byte for byte, what one rank computes alone for the same slices.  Uses volume 1, which differs per rank; rank 0
    rebuilds every rank's slices (they are a pure function of the rank)."""
    import torch
    import torch.distributed as dist

    from protosam_b200.engine import CoarseVolumeEngine
    cfg, world, rank = pipe.cfg, pipe.world, pipe.rank
    pipe.sync_all()
    out = pipe.step(0, lane=0, eager=True, vol_index=1)
    with torch.cuda.stream(pipe.lanes[0]):
        hdr_all, recs_all = out.result()
    pipe.pending[0] = None
    pipe.sync_all()
    res = None
    if rank == 0:
        solo = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], out_size=1024, val_wsize=cfg["ws"], use_cca=False,
                                  point_mode="both", match_algo=args.algo)
        solo.set_support(pipe.sup, pipe.fg, broadcast=False)
        ok, n_img, n_rec = True, 0, 0
        Hs = np.frombuffer(hdr_all.cpu().numpy().tobytes(), dtype=_HDR_SPLIT)
        R = recs_all.cpu().numpy()
        for r in range(world):
            h1, r1 = solo.run(pipe.rank_volume(1, r))
            torch.cuda.synchronize()
            h1n = np.frombuffer(h1.cpu().numpy().tobytes(), dtype=_HDR_SPLIT)
            r1n = r1.cpu().numpy()
            cnt = pipe.counts_all[r]
            blk = Hs[n_img: n_img + cnt]
            for f in ("ncc", "n_rec", "body"):                               # every header field but the record offset
                ok &= bool(np.array_equal(blk[f], h1n[f]))
            for i in range(cnt):
                k, a = int(h1n["n_rec"][i]), int(blk["first"][i])
                ok &= bool(np.array_equal(R[a: a + k], r1n[i, :k]))
                n_rec += k
            n_img += cnt
        res = {"ranks": world, "images": int(n_img), "records": int(n_rec), "ok": bool(ok),
               "what": "records gathered over NCCL (sharded slices, broadcast prototypes, compact gather) == single-rank "
                       "run of the same slices, byte for byte"}
    if world > 1:
        dist.barrier()
    return res


def gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist

    from protosam_b200 import _lib, ops, synth
    from protosam_b200.engine import GraphedVolumeStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / warnings off stdout: one JSON line only
        if args.nccl_channels > 0:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.nccl_channels))   # small messages: few CTAs suffice
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    # N > 1: the persistent GEMM grid leaves a few SMs to short-lived CTAs, so that the NCCL kernels of the volumes in
    # flight do not wait for a GEMM CTA to retire (psam_match_reserve_sms; 2 GPUs: 0.346 -> 0.328 ms per volume with 8)
    reserve = args.reserve_sms if args.reserve_sms >= 0 else (8 if world > 1 else 0)
    _lib.match_reserve_sms(reserve)

    Q, L, C, h, w = cfg["Q"], cfg["L"], cfg["C"], cfg["h"], cfg["w"]
    NL = max(1, args.lanes)

    # one process group (= NCCL communicator + stream) per lane: the collectives of different lanes must not
    # queue behind each other
    def lane_group():
        # high-priority NCCL stream: the small collective kernels must not queue behind the pending CTAs of the
        # big compute kernels of other lanes (every rank would wait for the slowest one)
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            return dist.new_group(list(range(world)), pg_options=opts)
        except Exception:
            return dist.new_group(list(range(world)))
    groups = [lane_group() if world > 1 else None for _ in range(NL)]

    scaling = args.scaling
    q_total = Q * world if scaling == "weak" else Q
    counts_all = shard_counts(q_total, world, L)
    q_local = counts_all[rank] // L
    pipe = Pipeline(cfg, q_local, q_total, counts_all, args, world, rank, dev, groups, N_ROTATE, not args.no_graphs)
    eng, lanes, qvols = pipe.engs[0], pipe.lanes, pipe.qvols

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        pipe.step(i)
    pipe.sync_all()
    line0 = sampler.mark() if rank == 0 else 0

    K = args.steps
    n0 = _lib.launch_count()
    pipe.replayed = 0
    ms_step, host_ms, evs = pipe.timed(K, 0)
    launches = _lib.launch_count() - n0 + pipe.replayed
    if pipe.graphs is None:
        ms_match = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
        ms_prompt = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
        ms_volume = None
    else:           # graph replays: only the whole volume is bracketed (the per-kernel pass below splits it)
        ms_match = ms_prompt = None
        ms_volume = float(np.mean([e[0].elapsed_time(e[2]) for e in evs]))
    # keep the same kernels running briefly if the timed region was too short for clock samples (rank-local work
    # only: no collective may be entered by a subset of the ranks)
    if rank == 0 and sampler.mark() - line0 < 3:
        t_end_probe = time.time() + 1.0
        i = 0
        while time.time() < t_end_probe:
            eng.run(qvols[i % N_ROTATE]); i += 1
            if i % 20 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.stop(line0) if rank == 0 else None

    # ---- per-kernel device times: the same step with the library's event pairs around every launch ------
    Kp = max(3, min(K, 20))
    _lib.profile_collect()
    _lib.profile_enable(True)
    for i in range(Kp):
        out_p = pipe.step(i, lane=0, eager=True)   # one lane, no graph replay: kernels run back to back, so each event pair times one kernel
    if world == 1:
        hdr_p, recs_p = out_p
    else:
        with torch.cuda.stream(lanes[0]):
            hdr_p, recs_p = out_p.result()
    pipe.pending[0] = None
    pipe.sync_all()
    _lib.profile_enable(False)
    prof = {k: (ms / Kp, n // Kp) for k, (ms, n) in _lib.profile_collect().items()}   # ms per step, launches per step
    H = ops.decode_headers(hdr_p) if (rank == 0 and hdr_p is not None) else None

    # ---- N > 1: what rank 0 gathered == a single-rank run ---------------------------------------------
    parity = parity_check(pipe, args) if world > 1 else None

    # ---- end-to-end through the public API from pinned host buffers ---------------------------------------------
    # Every step copies its own inputs host->device and its COMPACT records device->host inside the timed region;
    # with NL lanes the copies of one volume overlap the kernels of the previous one, and the host waits
    # for a volume's records before that lane's buffers are reused (and for all of them at the end).  At N > 1 every
    # rank reads its own records back (no gather in this leg).
    Ke = max(3, min(K, 50))
    vol = pipe.vol
    sup, fg, base = pipe.sup, pipe.fg, pipe.base
    h_sup = torch.from_numpy(vol.sup).pin_memory()
    h_fg = torch.from_numpy(vol.fg).pin_memory()
    h_q = [q.cpu().pin_memory() for q in qvols]
    n_img = q_local * L
    n_alloc = max(counts_all)
    cap = n_alloc * eng.recs_per_image
    nb_packed = ops.packed_bytes(n_alloc, cap)
    lane_buf = []
    for ln in range(NL):
        with torch.cuda.stream(lanes[ln]):
            lane_buf.append(dict(d_sup=torch.empty_like(sup), d_fg=torch.empty_like(fg), d_q=torch.empty_like(base),
                                 h_packed=torch.empty(nb_packed, dtype=torch.uint8).pin_memory(),
                                 done=torch.cuda.Event()))
    pipe.sync_all()
    busy = [False] * NL
    e2e_graphs = None
    if pipe.graphs is not None:
        e2e_graphs = []
        for ln in range(NL):
            with torch.cuda.stream(lanes[ln]):
                B = lane_buf[ln]
                B["d_sup"].copy_(sup); B["d_fg"].copy_(fg); B["d_q"].copy_(base)
                e2e_graphs.append(GraphedVolumeStep(pipe.engs[ln], B["d_sup"], B["d_fg"], B["d_q"], q_total=q_total,
                                                    split_streams=pipe.split))
        pipe.sync_all()

    def e2e_step(i):
        ln = i % NL
        B, e = lane_buf[ln], pipe.engs[ln]
        if busy[ln]:
            B["done"].synchronize()                         # the caller consumes that volume's prompts on the host
        with torch.cuda.stream(lanes[ln]):
            B["d_sup"].copy_(h_sup, non_blocking=True)
            B["d_fg"].copy_(h_fg, non_blocking=True)
            B["d_q"].copy_(h_q[i % N_ROTATE], non_blocking=True)
            if e2e_graphs is not None:
                e2e_graphs[ln].launch(gather=False)
                e2e_graphs[ln].join()
                packed = e2e_graphs[ln].buf
            else:
                e.set_support(B["d_sup"], B["d_fg"], broadcast=False)
                _, _, packed = e.prompts_from_logits(e.match(B["d_q"]), n_alloc=n_alloc, return_packed=True)
            B["h_packed"].copy_(packed, non_blocking=True)
            B["done"].record()
        busy[ln] = True

    def e2e_drain():
        for ln in range(NL):
            if busy[ln]:
                lane_buf[ln]["done"].synchronize()
                busy[ln] = False

    for i in range(3):
        e2e_step(i)
    e2e_drain()
    pipe.sync_all()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_step(i)
    e2e_drain()
    pipe.sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / Ke
    te = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    h2d = h_sup.numel() * 4 + h_fg.numel() * 4 + h_q[0].numel() * 4
    d2h = lane_buf[0]["h_packed"].numel()
    e2e_ok = None
    if rank == 0:       # the records that reached the host are complete and decodable
        Hh, Rh = ops.decode_packed(lane_buf[(Ke - 1) % NL]["h_packed"], n_alloc, cap, n_img)
        e2e_ok = bool(len(Hh) == n_img and int(Hh["n_rec"].sum()) == len(Rh))
    counts = eng.protos["counts"].cpu().numpy()
    del lane_buf, e2e_graphs, h_q

    # ---- the sharded runs BASELINE.json names (strong scaling: fixed totals) ------------------------------------
    extra = {}
    if not args.no_north_star_runs and scaling == "weak" and cfg["name"] == WORKLOAD:
        del pipe, eng, qvols
        torch.cuda.empty_cache()
        for name, steps_x in (("cfg3_synapse_ct", 6), ("cfg5_stress_vitl", 4)):
            cx = dict(synth.CONFIGS[name]); cx["name"] = name; cx["lanes"] = 2
            ax = argparse.Namespace(**vars(args)); ax.lanes = 2
            cnt = shard_counts(cx["Q"], world, cx["L"])
            px = Pipeline(cx, cnt[rank] // cx["L"], cx["Q"], cnt, ax, world, rank, dev,
                          groups[:2] if (world > 1 and NL >= 2) else [groups[0], groups[0]], 1, not args.no_graphs)
            ms_x, _, _ = px.timed(steps_x, 2)
            extra[name] = {"scaling": "strong", "slices_total": cx["Q"], "labels": cx["L"], "slices_per_gpu": cnt[rank] // cx["L"],
                           "value": cx["Q"] / (ms_x * 1e-3), "unit": UNIT, "ms_per_step": ms_x, "steps": steps_x, "warmup": 2,
                           "features": f"{cx['h']}x{cx['w']}x{cx['C']}", "n_gpus": world,
                           "l2_policy": "one volume per rank, >= 150 MB of query features per rank and step (> 126 MB L2)"}
            del px
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_tab = {}
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    sumP = int(counts.sum())
    out = 1024
    n_fg = int(H["n_fg"][: n_img].sum()) if H is not None else 0
    n_cc = int(H["ncc"][: n_img].sum()) if H is not None else 0
    flops_match = 2.0 * q_local * h * w * C * sumP             # algorithmic: 2*HW*C*sum(P) per slice
    model = {   # kernel -> (bound, algorithmic work per launch, note)
        "k_match_ts": ("tensor", flops_match, "2*HW*C*sum(P) per slice; executed = 3x (split-bf16 passes); fp32 query "
                       "converted inside the GEMM, A operand in TMEM"),
        "k_match_tc": ("tensor", flops_match, "2*HW*C*sum(P) per slice; executed = 3x (split-bf16 passes)"),
        "k_match_simt": ("tensor", flops_match, "2*HW*C*sum(P) per slice on CUDA cores"),
        "k_blocks_warp": ("issue", n_img * (8.0 * h * w + out * out / 8.0) + 4.0 * n_fg,
                          "CUDA-core ISSUE-bound (exact ATen-CPU arithmetic per pixel): achieved = issue slots busy, from the "
                          "ncu capture in profiles/ (smsp__issue_active); algorithmic bytes per image for reference: 8*h*w "
                          "logits + out^2/8 mask bits + 4*n_fg probabilities"),
        "k_components": ("hbm", n_img * (out * out / 8.0 + 64.0) + 4.0 * n_fg + 96.0 * n_cc,
                         "per image: out^2/8 mask bits + 4*n_fg probabilities + 96 B per component (latency-bound integer work)"),
        "k_pack_query": ("hbm", 8.0 * q_local * h * w * C, "4*C*HW read + 4*C*HW operand image written per slice"),
    }
    dom = max(prof, key=lambda k: prof[k][0]) if prof else "k_match_ts"
    ms_dom, n_dom = prof.get(dom, (ms_match, 1))
    n_dom = max(n_dom, 1)
    bound, work, note = model.get(dom, ("hbm", 0.0, "no model"))
    sec = ms_dom / n_dom * 1e-3
    if bound == "tensor":
        ach, peak, unit = work / sec / 1e12, peak_tf, "TFLOP/s"
    elif bound == "issue":
        ach, peak, unit = float(traffic_tab.get("issue_slot_busy_pct", {}).get(dom, 0.0)), 100.0, "% of issue slots (ncu)"
    else:
        ach, peak, unit = work / sec / 1e9, peak_bw, "GB/s"
    roof = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
            "traffic": traffic_tab.get(dom), "peak_source": peak_src, "ms_per_launch": ms_dom / n_dom,
            "share_of_step": ms_dom / sum(v[0] for v in prof.values()) if prof else None,
            "algorithmic_work_per_launch": work, "work_model": note, "sum_prototypes": sumP,
            "kernels_ms_per_step": {k: round(v[0], 5) for k, v in prof.items()},
            "match_stage_ms": ms_match, "prompt_stage_ms": ms_prompt, "volume_latency_ms": ms_volume,
            "issue_slot_busy_pct": traffic_tab.get("issue_slot_busy_pct", {})}
    gemm = next((k for k in ("k_match_ts", "k_match_tc") if k in prof), None)
    if gemm:
        t = prof[gemm][0] * 1e-3
        mk = {"kernel": gemm, "bound": "tensor", "achieved": flops_match / t / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
              "frac": flops_match / t / 1e12 / peak_tf, "executed_frac": 3 * flops_match / t / 1e12 / peak_tf,
              "ceiling_frac": 1.0 / 3.0, "traffic": traffic_tab.get(gemm), "ms_per_launch": prof[gemm][0]}
        if dom == gemm:
            roof["executed_frac"], roof["ceiling_frac"] = mk["executed_frac"], mk["ceiling_frac"]
        else:
            roof["match_kernel"] = mk
        # the fused ALP path of north_star = kernels 1 + 2 (prototypes, operand images, contraction + softmax epilogue)
        alp = sum(v[0] for k, v in prof.items() if k.startswith(("k_proto_stage", "k_pack_", "k_match_")))
        roof["alp_path"] = {"kernels": " + ".join(k for k in prof if k.startswith(("k_proto_stage", "k_pack_", "k_match_"))),
                            "ms_per_step": alp,
                            "achieved": flops_match / (alp * 1e-3) / 1e12, "unit": "TFLOP/s",
                            "frac": flops_match / (alp * 1e-3) / 1e12 / peak_tf, "ceiling_frac": 1.0 / 3.0}

    # ---- CPU baseline: the oracle port, one core, bounded sample ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ns = 12                                       # ~10 s of one host core
        val, dt = run_cpu_port(cfg, ns, 1, 1, 0)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{ns} query slices x {L} labels, one pass, C port of the reference path (oracle/), "
                         f"{dt:.1f} s"}

    line = {
        "metric": METRIC, "value": q_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload_desc(cfg, world, scaling, q_total, p2p=bool(args.p2p) and world > 1),
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": q_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "records_complete": e2e_ok,
                "note": "per rank: features + masks host->device, compact records device->host" +
                        ("; no gather in this leg (every rank reads its own records)" if world > 1 else "")},
        "gpu_launches": int(launches), "clocks": clocks, "host_enqueue_ms_per_step": host_ms,
        "parity_check": parity, "north_star_runs": extra or None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", type=int, default=0, help="match kernel: 0 auto, 1 fp32 CUDA cores, 2 tcgen05 (packed operands), 3 tcgen05 (fused conversion)")
    ap.add_argument("--slices", type=int, default=0, help="override the workload's query slices (per GPU when weak, in total when strong)")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's slices per GPU; strong: the workload's slices in total, sharded over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star-runs", action="store_true", help="skip the config-3 / config-5 sharded runs appended to the line")
    ap.add_argument("--port-only", action="store_true", help="--impl reference: time the C port even where the reference tree is mounted")
    ap.add_argument("--nccl-channels", type=int, default=4,
                    help="cap NCCL channels (0 = NCCL's default): the collectives move a few MB, and every extra channel "
                         "is a CTA that competes with the compute kernels for SMs (4 measured best at 2 GPUs in round 2: "
                         "0.382 ms/step vs 0.406 with 2 and 0.390 with 8)")
    ap.add_argument("--p2p", type=int, default=1,
                    help="N > 1: 1 = prototype table and prompt records move with one-sided stores over NVLink peer memory "
                         "(psam_peer_*, no collective library on the path); 0 = NCCL broadcast + gather")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs the persistent match kernel leaves free (psam_match_reserve_sms); -1 = 8 when N > 1, else 0")
    ap.add_argument("--no-graphs", action="store_true", help="enqueue every kernel from Python instead of CUDA graphs")
    ap.add_argument("--lanes", type=int, default=4, help="volumes in flight per GPU (CUDA streams)")
    ap.add_argument("--graph-collectives", type=int, default=0,
                    help="1 (N > 1): capture the NCCL broadcast and gather into the volume's CUDA graph (one launch per volume)")
    ap.add_argument("--split-streams", type=int, default=0,
                    help="1: replay a volume's prompt stage on a second, higher-priority stream (engine.GraphedVolumeStep)")
    args = ap.parse_args()
    from protosam_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    cfg["name"] = args.workload
    if args.slices:
        cfg["Q"] = args.slices
    cfg["lanes"] = args.lanes
    if args.impl == "reference":
        reference_arm(args, cfg)
    else:
        gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
