#!/usr/bin/env python
"""bench.py -- query slices/sec through the ALP match + prompt path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU port of the reference path

One step = one synthetic volume of the CHAOS-MRI-shaped config (BASELINE.json configs[1]): kernel 1
over the support slice for 4 labels (8 prototype sets), kernel 2 over 32 query slices, kernel 3 over
the 128 resulting coarse maps -> prompt records.  With N ranks every rank takes 32 further slices of
the same volume (weak scaling): one rank computes the prototypes and broadcasts them (NCCL), every rank
matches its slices, prompt records are gathered to rank 0.  `--lanes` volumes are in flight per GPU on
separate CUDA streams, each replayed from CUDA graphs (engine.GraphedVolumeStep).

`value` is timed with inputs resident in HBM; `e2e` goes through the same engine from pinned host
buffers with the host->device copies and the device->host read of the records inside the timed
region.  The JSON line also carries the roofline of the dominant kernel (timed live with CUDA
events around its launch) and a CPU baseline (the oracle port, one core, bounded sample).
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
N_ROTATE = 4            # distinct input volumes cycled through the timed region (> L2 in total)


def workload_desc(cfg, n_gpus):
    return {
        "workload": (f"{cfg['name']}: 1 support + {cfg['Q']} query slices per GPU, /14 patch features "
                     f"{cfg['h']}x{cfg['w']}x{cfg['C']}, {cfg['L']} labels (bg 'gridconv' + fg 'gridconv+'/'mask' "
                     f"decided on device), ws={cfg['ws']}, upsample {cfg['img_size']}->1024, prompts for every "
                     f"component (use_cca=False, point_mode=both)"),
        "slices_per_step_per_gpu": cfg["Q"],
        "labels": cfg["L"],
        "l2_policy": f"{N_ROTATE} distinct query volumes rotate through the timed region "
                     f"({N_ROTATE * cfg['Q'] * cfg['h'] * cfg['w'] * cfg['C'] * 4 / 1e6:.0f} MB of inputs + "
                     f"{cfg['Q'] * cfg['L'] * 4.2:.0f} MB of per-step intermediates > 126 MB L2)",
        "parallelism": f"slices sharded over {n_gpus} GPU(s); prototype broadcast (source rank = lane % ranks) + "
                       "record gather to rank 0",
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
# CPU port (oracle) of the same step -- cpu_baseline and the --impl reference arm
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


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_slices = max(1, min(cfg["Q"], cores))
    # keep the whole run within a few minutes: one slice (4 labels) costs ~2 s of one core
    steps = max(1, min(args.steps, 5))
    warm = max(0, min(args.warmup, 1))
    val, dt = run_cpu_port(cfg, n_slices, cores, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_desc(cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_slices} query slices x {cfg['L']} labels per step on {cores} threads "
                                   "(C port of the reference path: the reference is Python and is not on this box)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist

    from protosam_b200 import _lib, ops, synth
    from protosam_b200.engine import CoarseVolumeEngine, GraphedVolumeStep, gather_packed, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / warnings off stdout: one JSON line only
        if args.nccl_channels > 0:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.nccl_channels))   # MB-sized messages: few CTAs suffice
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    Q, L, C, h, w = cfg["Q"], cfg["L"], cfg["C"], cfg["h"], cfg["w"]
    # synthetic inputs: one support slice; N_ROTATE query volumes per rank (different seeds per rank)
    vol = synth.make_volume(1234, Q=Q, L=L, C=C, h=h, w=w, img_size=cfg["img_size"])
    sup = torch.from_numpy(vol.sup).to(dev)
    fg = torch.from_numpy(vol.fg).to(dev)
    base = torch.from_numpy(vol.qry).to(dev)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    qvols = [base] + [(base.roll(k + 1, 0) + 0.05 * torch.randn(base.shape, generator=g, device=dev)).contiguous()
                      for k in range(N_ROTATE - 1)]
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
    engs = [CoarseVolumeEngine((h, w), cfg["img_size"], out_size=1024, val_wsize=cfg["ws"], use_cca=False,
                               point_mode="both", match_algo=args.algo, group=groups[ln]) for ln in range(NL)]
    eng = engs[0]
    lanes = [torch.cuda.Stream(device=dev) for _ in range(NL)]
    q_total = Q * world
    counts_all = [(b - a) * L for a, b in (shard_range(q_total, world, r) for r in range(world))]
    pending = [None] * NL
    replayed = [0]          # library kernels launched through graph replays
    # CUDA graphs: one per (lane, resident volume); volume i goes to lane i % NL, graph (i // NL) % len(graphs[lane])
    graphs = None
    if not args.no_graphs:
        graphs = []
        for ln in range(NL):
            with torch.cuda.stream(lanes[ln]):
                mine = [qvols[j] for j in range(N_ROTATE) if j % NL == ln] or [qvols[ln % N_ROTATE]]
                graphs.append([GraphedVolumeStep(engs[ln], sup, fg, qv, q_total=q_total, src=ln % world) for qv in mine])
        torch.cuda.synchronize()

    def step(i, ev=None, lane=None, eager=False):
        """One volume: prototypes (+ broadcast), match, prompts, record gather, all on the stream of lane
        i % NL: consecutive volumes are in flight on different streams, so the small launches, the collectives
        and the tail of one volume overlap the big kernels of the next.  The gather of a volume is asynchronous
        (NCCL's stream) and is waited for when its lane is used again / at the end of the region."""
        ln = i % NL if lane is None else lane
        e = engs[ln]
        with torch.cuda.stream(lanes[ln]):
            if pending[ln] is not None:
                pending[ln].wait()
            if graphs is not None and not eager:
                gs = graphs[ln][(i // NL) % len(graphs[ln])]
                if ev is not None:
                    ev[0].record()
                out = gs.launch()
                replayed[0] += gs.n_kernels
                if ev is not None:
                    ev[2].record()
                if world > 1:
                    pending[ln] = out
                return out
            e.set_support(sup, fg, src=ln % world)
            qv = qvols[i % N_ROTATE]
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
            _, _, buf = e.prompts_from_logits(logits, n_alloc=max(counts_all), return_packed=True)
            if ev is not None:
                ev[2].record()
            pending[ln] = gather_packed(buf, counts_all, e.max_cc, dst=0, group=groups[ln], async_op=True)
            return pending[ln]

    def finish(out, lane=0):
        """(hdr, recs) of a step's return value (waits for its gather on the lane's stream)"""
        if world == 1:
            return out
        with torch.cuda.stream(lanes[lane]):
            return out.result()

    def drain():
        """all lanes: wait for the pending gathers, then make the current stream wait for the lanes"""
        for ln in range(NL):
            with torch.cuda.stream(lanes[ln]):
                if pending[ln] is not None:
                    pending[ln].wait()
                    pending[ln] = None
            torch.cuda.current_stream().wait_stream(lanes[ln])

    def sync_all():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    line0 = sampler.mark() if rank == 0 else 0

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    replayed[0] = 0
    sync_all()
    t_start.record()
    for ln in range(NL):
        lanes[ln].wait_event(t_start)
    th0 = time.perf_counter()
    for i in range(K):
        step(i, evs[i])
    host_ms = (time.perf_counter() - th0) * 1e3 / K     # host time to enqueue one step (launch-bound if ~ ms_per_step)
    drain()                             # the last gathers are part of the timed region
    t_end.record()
    sync_all()
    launches = _lib.launch_count() - n0 + replayed[0]
    ms_total = t_start.elapsed_time(t_end)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    if graphs is None:
        ms_match = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
        ms_prompt = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
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
        out_p = step(i, lane=0, eager=True)   # one lane, no graph replay: kernels run back to back, so each event pair times one kernel
    hdr_p, recs_p = finish(out_p)
    pending[0] = None
    sync_all()
    _lib.profile_enable(False)
    prof = {k: (ms / Kp, n // Kp) for k, (ms, n) in _lib.profile_collect().items()}   # ms per step, launches per step
    H = ops.decode_headers(hdr_p) if (rank == 0 and hdr_p is not None) else None

    # ---- end-to-end through the public API from pinned host buffers ---------------------------
    # Every step copies its own inputs host->device and its records device->host inside the timed region;
    # with NL lanes the copies of one volume overlap the kernels of the previous one, and the host waits
    # for a volume's records before that lane's buffers are reused (and for all of them at the end).
    Ke = max(3, min(K, 50))
    h_sup = torch.from_numpy(vol.sup).pin_memory()
    h_fg = torch.from_numpy(vol.fg).pin_memory()
    h_q = [q.cpu().pin_memory() for q in qvols]
    n_img = Q * L
    lane_buf = []
    for ln in range(NL):
        with torch.cuda.stream(lanes[ln]):
            lane_buf.append(dict(d_sup=torch.empty_like(sup), d_fg=torch.empty_like(fg), d_q=torch.empty_like(base),
                                 h_hdr=torch.empty((n_img, 64), dtype=torch.uint8).pin_memory(),
                                 h_rec=torch.empty((n_img, eng.max_cc, 96), dtype=torch.uint8).pin_memory(),
                                 done=torch.cuda.Event()))
    sync_all()
    busy = [False] * NL
    e2e_graphs = None
    if graphs is not None:
        e2e_graphs = []
        for ln in range(NL):
            with torch.cuda.stream(lanes[ln]):
                B = lane_buf[ln]
                B["d_sup"].copy_(sup); B["d_fg"].copy_(fg); B["d_q"].copy_(base)
                e2e_graphs.append(GraphedVolumeStep(engs[ln], B["d_sup"], B["d_fg"], B["d_q"], q_total=q_total))
        sync_all()

    def e2e_step(i):
        ln = i % NL
        B, e = lane_buf[ln], engs[ln]
        if busy[ln]:
            B["done"].synchronize()                         # the caller consumes that volume's prompts on the host
        with torch.cuda.stream(lanes[ln]):
            B["d_sup"].copy_(h_sup, non_blocking=True)
            B["d_fg"].copy_(h_fg, non_blocking=True)
            B["d_q"].copy_(h_q[i % N_ROTATE], non_blocking=True)
            if e2e_graphs is not None:
                hd, rc = e2e_graphs[ln].launch(gather=False)
            else:
                e.set_support(B["d_sup"], B["d_fg"])
                hd, rc = e.run(B["d_q"])
            B["h_hdr"].copy_(hd, non_blocking=True)
            B["h_rec"].copy_(rc, non_blocking=True)
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
    sync_all()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_step(i)
    e2e_drain()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / Ke
    te = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    h2d = h_sup.numel() * 4 + h_fg.numel() * 4 + h_q[0].numel() * 4
    d2h = lane_buf[0]["h_hdr"].numel() + lane_buf[0]["h_rec"].numel()

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
    counts = eng.protos["counts"].cpu().numpy()
    sumP = int(counts.sum())
    n_img, out = Q * L, 1024
    n_fg = int(H["n_fg"][: n_img].sum()) if H is not None else 0
    n_cc = int(H["ncc"][: n_img].sum()) if H is not None else 0
    flops_match = 2.0 * Q * h * w * C * sumP                 # algorithmic: 2*HW*C*sum(P) per slice
    bytes_match = 4.0 * Q * h * w * C + 4.0 * C * sumP + 4.0 * Q * 2 * L * h * w
    model = {   # kernel -> (bound, algorithmic work per launch, note)
        "k_match_tc": ("tensor", flops_match, "2*HW*C*sum(P) per slice; executed = 3x (split-bf16 passes)"),
        "k_match_simt": ("tensor", flops_match, "2*HW*C*sum(P) per slice on CUDA cores"),
        "k_blocks_warp": ("hbm", n_img * (8.0 * h * w + out * out / 8.0) + 4.0 * n_fg,
                          "per image: 8*h*w logits + out^2/8 mask bits + 4*n_fg probabilities (upper bound: p_fg is "
                          "written only where kernel 3b reads it); the kernel is CUDA-core issue-bound, see profiles/"),
        "k_components": ("hbm", n_img * (out * out / 8.0 + 64.0) + 4.0 * n_fg + 96.0 * n_cc,
                         "per image: out^2/8 mask bits + 4*n_fg probabilities + 96 B per component"),
        "k_pack_query": ("hbm", 8.0 * Q * h * w * C, "4*C*HW read + 4*C*HW operand image written per slice"),
    }
    dom = max(prof, key=lambda k: prof[k][0]) if prof else "k_match_tc"
    ms_dom, n_dom = prof.get(dom, (ms_match, 1))
    n_dom = max(n_dom, 1)
    bound, work, note = model.get(dom, ("hbm", 0.0, "no model"))
    sec = ms_dom / n_dom * 1e-3
    if bound == "tensor":
        ach, peak, unit = work / sec / 1e12, peak_tf, "TFLOP/s"
    else:
        ach, peak, unit = work / sec / 1e9, peak_bw, "GB/s"
    roof = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
            "traffic": traffic_tab.get(dom), "peak_source": peak_src, "ms_per_launch": ms_dom / n_dom,
            "share_of_step": ms_dom / sum(v[0] for v in prof.values()) if prof else None,
            "algorithmic_work_per_launch": work, "work_model": note, "sum_prototypes": sumP,
            "kernels_ms_per_step": {k: round(v[0], 5) for k, v in prof.items()},
            "match_stage_ms": ms_match, "prompt_stage_ms": ms_prompt,
            "volume_latency_ms": ms_volume if graphs is not None else None}
    if dom != "k_match_tc" and "k_match_tc" in prof:
        t = prof["k_match_tc"][0] * 1e-3
        roof["match_kernel"] = {"kernel": "k_match_tc", "bound": "tensor", "achieved": flops_match / t / 1e12,
                                "peak": peak_tf, "unit": "TFLOP/s", "frac": flops_match / t / 1e12 / peak_tf,
                                "executed_frac": 3 * flops_match / t / 1e12 / peak_tf,
                                "traffic": traffic_tab.get("k_match_tc"), "ms_per_launch": prof["k_match_tc"][0]}
    elif dom == "k_match_tc":
        roof["executed_frac"] = 3 * roof["frac"]

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
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_desc(cfg, world),
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": q_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches), "clocks": clocks, "host_enqueue_ms_per_step": host_ms,
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
    ap.add_argument("--slices", type=int, default=0, help="override query slices per GPU")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-channels", type=int, default=2,
                    help="cap NCCL channels (0 = NCCL's default): the collectives move a few MB, and every extra channel "
                         "is a CTA that competes with the compute kernels for SMs (2 measured best at 4 and 8 GPUs)")
    ap.add_argument("--no-graphs", action="store_true", help="enqueue every kernel from Python instead of CUDA graphs")
    ap.add_argument("--lanes", type=int, default=4, help="volumes in flight per GPU (CUDA streams)")
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
