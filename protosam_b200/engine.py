"""Batched volume driver: support prototypes once per volume, then every query slice x label in
three launches, sharded over GPUs.

The reference processes one query slice and one label per Python iteration
(validation_protosam.py:352-388) and recomputes the same support prototypes for every slice
(grid_proto_fewshot.py:181-184, 239-259).  Here:

  set_support()  kernel 1 once per volume for the 2L prototype sets (bg_l 'gridconv', fg_l decided on
                 the device like grid_proto_fewshot.py:250-256) -- on rank 0, then ONE broadcast
  run()          kernel 2 over the rank's query slices against all sets (one launch), kernel 3a/3b
                 over the resulting [Q_local*L, 2, h, w] coarse maps (two launches)
  gather         ONE gather of the fixed-size prompt headers/records

Query slices are independent given the prototypes, so the only exchanges are that broadcast and
that gather.  No host synchronisation happens between set_support() and the
final device->host read of the records.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import ops, prompts as P

FG_THRESH = BG_THRESH = 0.95      # models/grid_proto_fewshot.py:21-22


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n items owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_PROTO_KEYS = ("protos", "counts", "eff_modes", "status")


def broadcast_prototypes(protos: dict, src: int = 0, group=None) -> dict:
    """Broadcast the prototype table from `src` (NCCL over NVLink on the GPU box; the same code runs over
    gloo in the CPU tests).  Shapes are rank-independent, so receivers pre-allocate; tables built by
    ops.proto_table_alloc live in one buffer and move with ONE collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return protos
    if protos.get("packed") is not None:
        dist.broadcast(protos["packed"], src=src, group=group)
        return protos
    for k in _PROTO_KEYS:
        dist.broadcast(protos[k], src=src, group=group)
    return protos


class PendingGather:
    """Handle of an in-flight record gather (NCCL runs it on its own stream, so the next volume's kernels
    overlap it).  result() waits and returns (hdr_all, recs_all) on dst, (None, None) elsewhere.

    layout = ("dense", max_cc): per-rank buffers are ops.records_alloc images, the result is dense
             (hdr [n,64], recs [n,max_cc,96]);
             ("compact", n_alloc, capacity): per-rank buffers are ops.compact_records images, the result is compact
             (hdr [n,64] with hdr.reserved rebased to the merged record list, recs [total,96])."""

    def __init__(self, work, bucket, counts, layout, keep):
        self.work, self.bucket, self.counts, self.layout, self.keep = work, bucket, counts, layout, keep
        self._ready = None

    @classmethod
    def done(cls, pair):
        p = cls(None, None, [], ("dense", 0), None)
        p._ready = pair
        return p

    def wait(self):
        """make the current stream wait for the gather (no host block, no re-packing)"""
        if self.work is not None:
            self.work.wait()
            self.work = None

    def result(self):
        if self._ready is not None:
            return self._ready
        self.wait()
        if self.bucket is None:
            return None, None
        if self.layout[0] == "dense":
            parts = [ops.split_records(b, self.layout[1]) for b in self.bucket]
            return (torch.cat([h[:c] for (h, _), c in zip(parts, self.counts)], 0),
                    torch.cat([r[:c] for (_, r), c in zip(parts, self.counts)], 0))
        _, n_alloc, cap = self.layout
        parts = [ops.split_packed(b, n_alloc, cap) for b in self.bucket]
        tails = torch.stack([t for _, t, _ in parts]).cpu().numpy()          # one small device->host read
        T = tails.view(ops.TAIL_DTYPE).reshape(-1)
        if (T["flags"] & ops._lib.PACKED_OVERFLOW).any():
            raise RuntimeError("a rank produced more prompt records than the compact buffer holds; raise recs_per_image")
        hdrs, recs, base = [], [], 0
        for (h, _, r), c, t in zip(parts, self.counts, T):
            h = h[:c].clone()
            h.view(torch.int32)[:, 15] += base                                # hdr.reserved: first record of the image
            hdrs.append(h)
            recs.append(r[: int(t["total"])])
            base += int(t["total"])
        return torch.cat(hdrs, 0), torch.cat(recs, 0)


def gather_packed(buf: torch.Tensor, counts: Sequence[int], layout, dst: int = 0, group=None, async_op: bool = False):
    """ONE gather of the per-rank packed buffers (all of the same size: sized for max(counts) images).  `layout` as in
    PendingGather; an int is taken as ("dense", max_cc)."""
    if isinstance(layout, int):
        layout = ("dense", layout)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = list(torch.empty((world, buf.numel()), dtype=buf.dtype, device=buf.device).unbind(0)) if rank == dst else None
    work = dist.gather(buf, bucket, dst=dst, group=group, async_op=async_op)
    pend = PendingGather(work if async_op else None, bucket, list(counts), layout, buf)
    return pend if async_op else pend.result()


def gather_records(hdr: torch.Tensor, recs: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None):
    """Gather per-rank dense (hdr [n_r,64], recs [n_r,max_cc,96]) to `dst` in rank order.  `counts` = images
    per rank (known from shard_range, no size exchange needed).  Returns (hdr_all, recs_all) on dst,
    (None, None) elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hdr, recs
    nmax, max_cc = max(counts), recs.shape[1]
    buf, h, r = ops.records_alloc(nmax, max_cc, hdr.device)
    h[: hdr.shape[0]] = hdr
    r[: recs.shape[0]] = recs
    return gather_packed(buf, counts, ("dense", max_cc), dst=dst, group=group)


class CoarseVolumeEngine:
    """ALP match + prompt extraction for one volume on this rank's GPU."""

    def __init__(self, feature_hw: Sequence[int], img_size: int, out_size: int = 1024, val_wsize: int = 2,
                 proto_grid_size: int = 8, use_cca: bool = False, point_mode: str = "both",
                 max_cc: int = ops.DEFAULT_MAX_CC, max_runs: int = ops.DEFAULT_MAX_RUNS, fg_mode: str = "auto_fg",
                 match_algo: int = 0, group=None, variant: str = "protosam",
                 recs_per_image: int = ops.DEFAULT_RECS_PER_IMAGE, p2p: bool = False):
        self.h, self.w = int(feature_hw[0]), int(feature_hw[1])
        self.img_size, self.out_size = int(img_size), int(out_size)
        self.val_wsize = int(val_wsize)
        # MultiProtoAsConv.kernel_size (models/alpmodule.py:34): the window of the fg-mode decision
        self.kernel_size = (self.h // proto_grid_size, self.w // proto_grid_size)
        self.use_cca, self.point_mode = bool(use_cca), point_mode
        self.max_cc, self.max_runs = int(max_cc), int(max_runs)
        self.recs_per_image = int(recs_per_image)      # capacity of the compact record buffers, per image
        self.fg_mode, self.match_algo, self.group = fg_mode, match_algo, group
        # 'protosam': confidences from softmax(logits) (models/ProtoSAM.py:599-608); 'medsam': ProtoMedSAM hands cca()
        # probabilities, which it soft-maxes again (models/ProtoMedSAM.py:178-187) -- only boxes are used downstream
        if variant not in ("protosam", "medsam"):
            raise ValueError("variant must be 'protosam' or 'medsam'")
        self.variant = variant
        self.prob_mode = "softmax_twice" if variant == "medsam" else "softmax"
        self.protos: Optional[dict] = None
        self.n_labels, self.n_shots = 0, 1
        self._ws = None
        # p2p (world > 1, CUDA): the prototype table and the prompt records move with one-sided stores over NVLink peer
        # memory (ops.PeerChannel, psam_peer_*) instead of NCCL collectives; the channels are created on first use
        # (a collective, blocking step: the first set_support / run_sharded must not be inside a graph capture)
        self.p2p = bool(p2p)
        self._channels = {}
        self._retired = []

    # -- distributed helpers -------------------------------------------------------------------
    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def _channel(self, kind: str, payload_bytes: int, device) -> Optional["ops.PeerChannel"]:
        """The peer channel of this engine for `kind`, created on first use (collective).  If symmetric memory cannot be
        set up on ANY rank (driver / container without peer mapping), every rank falls back to the NCCL path: the
        ranks agree through one all-reduce, so nobody is left waiting in a one-sided protocol."""
        ch = self._channels.get(kind)
        if ch is None or ch.payload_bytes < payload_bytes:
            err = None
            try:
                ch = ops.PeerChannel(payload_bytes, self.group, device)
            except Exception as e:      # noqa: BLE001 -- whatever the allocator / rendezvous raises
                ch, err = None, e
            ok = torch.tensor([0 if ch is None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                import warnings
                warnings.warn(f"protosam_b200: peer-memory exchanges unavailable ({err!r}); using NCCL collectives")
                self.p2p = False
                return None
            if kind in self._channels:
                self._retired.append(self._channels[kind])       # CUDA graphs captured earlier still point into it
            self._channels[kind] = ch
        return ch

    def send_table(self, protos: dict, src: int = 0):
        """The prototype table from `src` to every rank: one-sided pushes of the live rows (p2p) or one NCCL / gloo
        broadcast of the packed table."""
        world, rank = self._world()
        if world == 1:
            return protos
        if not self.p2p:
            return broadcast_prototypes(protos, src=src, group=self.group)
        ch = self._channel("table", protos["packed"].numel(), protos["packed"].device)
        if ch is None:
            return broadcast_prototypes(protos, src=src, group=self.group)
        if rank == src:
            ops.peer_push_table(ch, protos)
        else:
            ops.peer_recv_table(ch, protos, src)
        return protos

    def collect_records(self, buf: torch.Tensor, counts: Sequence[int], layout, dst: int = 0, async_op: bool = False):
        """This rank's packed record buffer to `dst`: one-sided puts into dst's mailbox (p2p; dst copies the slots out and
        the returned handle reads that copy -- consume it before this engine's next collect_records) or one gather."""
        if not self.p2p:
            return gather_packed(buf, counts, layout, dst=dst, group=self.group, async_op=async_op)
        world, rank = self._world()
        slot = (buf.numel() + 15) // 16 * 16
        ch = self._channel("records", slot * world, buf.device)
        if ch is None:
            return gather_packed(buf, counts, layout, dst=dst, group=self.group, async_op=async_op)
        slot = ch.payload_bytes // world // 16 * 16          # the channel's slot size (>= this buffer)
        ops.peer_put(ch, buf, dst)
        bucket = None
        if rank == dst:
            out = torch.empty((world, slot), dtype=torch.uint8, device=buf.device)
            ops.peer_collect(ch, out)
            bucket = [o[: buf.numel()] for o in out.unbind(0)]
        pend = PendingGather(None, bucket, list(counts), layout, buf)
        return pend if async_op else pend.result()

    # -- support side ----------------------------------------------------------------------------
    def set_support(self, sup_feats: torch.Tensor, fg_masks: torch.Tensor, src: int = 0, broadcast: bool = True):
        """sup_feats [S,h,w,C] channels-last support features (S shots; the reference configs use 1);
        fg_masks [L,S,h,w] foreground masks at feature resolution (nearest-downsampled like
        grid_proto_fewshot.py:228-231).  Computes on `src`, broadcasts to the other ranks."""
        S, h, w, C = sup_feats.shape
        assert (h, w) == (self.h, self.w)
        L = fg_masks.shape[0]
        self.n_labels, self.n_shots = L, S
        world, rank = self._world()
        sup_x = sup_feats.permute(0, 3, 1, 2)                      # logical [S,C,h,w], channels-last storage
        fg = fg_masks.reshape(L, 1, S, h, w).to(torch.float32)
        if S == 1:
            sup_y = torch.cat([1.0 - fg, fg], dim=1).reshape(2 * L, S, h, w)   # (bg_0, fg_0, bg_1, fg_1, ...)
            modes, shots = ["gridconv", self.fg_mode] * L, None
        else:
            # several shots (grid_proto_fewshot.py:239-264): background from all shots at once, foreground once per
            # shot (that shot's features and mask only), element-wise max over the shots afterwards
            sup_y = torch.cat([1.0 - fg] + [fg] * S, dim=1).reshape(L * (1 + S), S, h, w)
            modes = (["gridconv"] + [self.fg_mode] * S) * L
            shots = ([-1] + list(range(S))) * L
        nsets = sup_y.shape[0]
        if world == 1 or rank == src:
            protos = ops.alp_prototypes(sup_x, sup_y, modes, (self.val_wsize, self.val_wsize), FG_THRESH,
                                        auto_ksize=self.kernel_size, shots=shots)
        else:
            gh, gw = h // self.val_wsize, w // self.val_wsize
            N = S * gh * gw
            protos = ops.proto_table_alloc(nsets, N + S, C, sup_feats.device)
            protos.update(N=N, gh=gh, gw=gw, S=S, C=C)
        self.protos = protos
        if broadcast:
            self.send_table(protos, src=src)
        return self.protos

    def set_support_from_image_masks(self, sup_feats: torch.Tensor, fg_img_masks: torch.Tensor, src: int = 0,
                                     broadcast: bool = True):
        """As set_support, from masks at image resolution [L,S,H,W]: the nearest-neighbour resize to the feature
        map that FewShotSeg.forward does first (grid_proto_fewshot.py:228-231) runs on the device too."""
        return self.set_support(sup_feats, ops.mask_nearest(fg_img_masks.to(torch.float32), self.h, self.w), src=src,
                                broadcast=broadcast)

    # -- query side ------------------------------------------------------------------------------
    def match(self, qry_feats: torch.Tensor) -> torch.Tensor:
        """qry_feats [Q,h,w,C] channels-last -> coarse logits [Q*L, 2, h, w] (bg, fg per label)."""
        Q, h, w, C = qry_feats.shape
        scores, _, _ = ops.alp_match(qry_feats.view(Q, h * w, C), self.protos, want_assign=False,
                                     algo=self.match_algo)
        if getattr(self, "n_shots", 1) > 1:
            return ops.combine_shots(scores, self.n_labels, self.n_shots).view(Q * self.n_labels, 2, h, w)
        return scores.view(Q * self.n_labels, 2, h, w)

    def prompts_from_logits(self, logits: torch.Tensor, n_alloc=None, return_packed=False):
        """coarse logits [n,2,h,w] -> (hdr uint8 [n,64], recs uint8 [n,max_cc,96]) on the device; return_packed adds the
        compact buffer (headers + live records only, sized for n_alloc images) that gathers and host copies move."""
        n = logits.shape[0]
        need = ops._lib.load().psam_coarse_to_prompts_workspace(n, self.out_size, self.max_runs, self.max_cc)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=logits.device)
        return ops.coarse_to_prompts(logits, self.img_size, self.out_size, self.use_cca, self.max_cc,
                                     self.max_runs, workspace=self._ws, n_alloc=n_alloc, return_packed=return_packed,
                                     prob_mode=self.prob_mode,
                                     capacity=max(n_alloc or 0, logits.shape[0]) * self.recs_per_image)

    def run(self, qry_feats: torch.Tensor):
        """-> (hdr, recs) on the device; image index = q*L + l."""
        return self.prompts_from_logits(self.match(qry_feats))

    def run_sharded(self, qry_local: torch.Tensor, q_total: int, dst: int = 0, async_op: bool = False):
        """This rank's block of a Q-slice volume (see shard_range) -> gathered records on `dst`: (hdr, recs), or a
        PendingGather when async_op (its result() gives the same pair)."""
        world, _ = self._world()
        counts = [(hi - lo) * self.n_labels for lo, hi in (shard_range(q_total, world, r) for r in range(world))]
        if world == 1:
            out = self.run(qry_local)
            return PendingGather.done(out) if async_op else out
        _, _, buf = self.prompts_from_logits(self.match(qry_local), n_alloc=max(counts), return_packed=True)
        return self.collect_records(buf, counts, ("compact", max(counts), max(counts) * self.recs_per_image), dst=dst,
                                    async_op=async_op)

    def decode(self, hdr: torch.Tensor, recs: torch.Tensor, on_empty_set: str = "raise") -> List[List[P.SlicePrompts]]:
        """Device records -> per slice, per label prompt objects (one D2H copy).  recs is dense [n,max_cc,96] or compact
        [total,96] (then image i owns recs[hdr.reserved : hdr.reserved + hdr.n_rec]).  on_empty_set: 'raise' (the
        reference's behaviour, see check_status) or 'ignore' (such labels decode as empty prompts: their scores are NaN)."""
        if on_empty_set == "raise":
            self.check_status()
        H = ops.decode_headers(hdr)
        if recs.dim() == 3:
            R = ops.decode_records(recs)
            per = [R[i] for i in range(len(H))]
        else:
            Rc = np.frombuffer(recs.detach().cpu().numpy().tobytes(), dtype=ops.REC_DTYPE)
            per = [Rc[int(h["reserved"]): int(h["reserved"]) + int(h["n_rec"])] for h in H]
        L = self.n_labels
        flat = [P.prompts_from_records(H[i], per[i], self.use_cca, self.point_mode) for i in range(len(H))]
        return [flat[q * L:(q + 1) * L] for q in range(len(flat) // L)]

    def check_status(self):
        """The reference raises when a 'gridconv' set has no prototype (F.conv2d on a [0,C,1,1] weight,
        models/alpmodule.py:68); the batched path records it in the device-side status words instead of synchronising
        per call.  decode() -- the first point where the host reads results anyway -- turns it back into the error."""
        if self.protos is not None and self.protos.get("status") is not None:
            st = self.protos["status"].detach().cpu().numpy()
            bad = np.nonzero(st & ops._lib.SET_EMPTY)[0]
            if len(bad):
                print("failed to find prototypes")
                raise RuntimeError(f"no prototypes survived the threshold in 'gridconv' set(s) {bad.tolist()} "
                                   "(the reference raises inside F.conv2d)")


class GraphedVolumeStep:
    """One volume's device work captured into CUDA graphs for fixed input buffers, so that a step costs the host
    two graph launches and two collectives instead of ~20 Python-level launches (the path is launch-bound on the
    host once several ranks share the host's cores):

        graph 1 (rank `src` only)  kernel 1 over the support slice -> prototype table
        broadcast                  of the table (one NCCL call, outside the graphs)
        graph 2                    kernel 2 + kernels 3a/3b over this rank's query slices -> packed records
        gather                     of the packed records (one NCCL call, asynchronous)

    `sup_feats`, `fg_masks`, `qry_local` are the buffers the graphs read: refill them in place (or copy new data
    into them on the same stream) before each `launch()`.  `self.buf` is the compact record buffer of the last
    launch (ops.decode_packed / ops.split_packed read it); it lives in graph 2's memory pool, so the next launch()
    rewrites it -- launch() therefore first makes the stream wait for the previous launch's asynchronous gather.
    """

    def __init__(self, eng: CoarseVolumeEngine, sup_feats: torch.Tensor, fg_masks: torch.Tensor,
                 qry_local: torch.Tensor, q_total: Optional[int] = None, src: int = 0, dst: int = 0,
                 split_streams: bool = False, capture_collectives: bool = False):
        """split_streams: graph 2 is cut in two -- the match stage stays on the caller's stream, the prompt stage
        (kernels 3a/3b, compaction, gather) is replayed on a second, HIGHER-priority stream of this object.  The GPU's
        work distributor hands out CTAs kernel by kernel in arrival order within a priority level: with several volumes
        in flight the GEMM grids of the other lanes (each needs every SM) queue up ahead of this volume's prompt kernels,
        which then wait although their CTAs would fit beside a resident GEMM CTA (measured with tools/trace_timeline.py:
        the lanes run phase-locked, four GEMMs back to back, then four prompt stages).  On a higher-priority stream the
        prompt kernels are dispatched as soon as they are ready, beside the GEMM of the next volume.  The next launch()'s
        match stage also no longer queues behind this volume's prompt stage.  join() makes the caller's stream wait
        for the prompt stage."""
        self.eng, self.src, self.dst = eng, src, dst
        self.split = bool(split_streams)
        # capture_collectives (world > 1): the broadcast and the gather are captured too, so a volume is ONE graph launch
        # per rank -- no host round trips between the kernels and the NCCL kernels, and no Python on the critical path
        self.one_graph = bool(capture_collectives)
        world, rank = eng._world()
        self.world, self.rank = world, rank
        L = fg_masks.shape[0]
        q_total = q_total if q_total is not None else qry_local.shape[0] * world
        self.counts = [(hi - lo) * L for lo, hi in (shard_range(q_total, world, r) for r in range(world))]
        n_alloc = max(self.counts)
        # eager warm-up on the current stream: sizes workspaces, sets kernel attributes, primes the allocator (and, with
        # the engine's p2p exchanges, creates the peer channels: a collective step that cannot be captured)
        p2p = bool(getattr(eng, "p2p", False)) and world > 1
        self.layout = ("compact", n_alloc, n_alloc * eng.recs_per_image)
        eng.set_support(sup_feats, fg_masks, src=src)
        _, _, buf0 = eng.prompts_from_logits(eng.match(qry_local), n_alloc=n_alloc, return_packed=True)
        if p2p:
            eng.collect_records(buf0, self.counts, self.layout, dst=dst)
            p2p = bool(eng.p2p)          # the engine falls back to NCCL when peer channels cannot be created
        torch.cuda.current_stream().synchronize()
        self.g1 = None
        k0 = ops._lib.launch_count()
        # p2p: the exchanges are kernels of this library, so a volume is always ONE graph (split_streams does not apply)
        self.split = self.split and not p2p
        self.one_graph = (self.one_graph or p2p) and world > 1 and not self.split
        self._p2p_pending = None
        if self.one_graph:
            nb = ops.packed_bytes(n_alloc, n_alloc * eng.recs_per_image)
            self.bucket = (list(torch.empty((world, nb), dtype=torch.uint8, device=qry_local.device).unbind(0))
                           if rank == dst and not p2p else None)
            self.gall = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.gall):
                eng.set_support(sup_feats, fg_masks, src=src)          # kernel 1 on `src` + the broadcast / push
                self.hdr, self.recs, self.buf = eng.prompts_from_logits(eng.match(qry_local), n_alloc=n_alloc,
                                                                        return_packed=True)
                if p2p:
                    self._p2p_pending = eng.collect_records(self.buf, self.counts, self.layout, dst=dst, async_op=True)
                else:
                    dist.gather(self.buf, self.bucket, dst=dst, group=eng.group)
            self.protos = eng.protos
            self.n_kernels = int(ops._lib.launch_count() - k0)
            self.n_alloc = n_alloc
            self._pending = None
            return
        if world == 1 or rank == src:
            self.g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g1):
                eng.set_support(sup_feats, fg_masks, src=src, broadcast=False)
        else:
            eng.set_support(sup_feats, fg_masks, src=src, broadcast=False)     # allocates the receive table
        self.protos = eng.protos
        self.g2 = torch.cuda.CUDAGraph()
        self.g3 = None
        if not self.split:
            with torch.cuda.graph(self.g2):
                self.hdr, self.recs, self.buf = eng.prompts_from_logits(eng.match(qry_local), n_alloc=n_alloc,
                                                                        return_packed=True)
        else:
            with torch.cuda.graph(self.g2):
                self.logits = eng.match(qry_local)
            self.g3 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g3):
                self.hdr, self.recs, self.buf = eng.prompts_from_logits(self.logits, n_alloc=n_alloc, return_packed=True)
            self.prompt_stream = torch.cuda.Stream(device=qry_local.device, priority=-1)
            self._match_done = torch.cuda.Event()
            self._prompts_done = torch.cuda.Event()
            self._prompts_recorded = False
        self.n_kernels = int(ops._lib.launch_count() - k0)      # library kernels one launch() replays
        self.n_alloc = n_alloc
        self._pending = None

    def launch(self, async_gather: bool = True, gather: bool = True):
        """Enqueue one volume on the current stream.  -> this rank's (hdr, recs) device views when world == 1 or
        gather is False, else a PendingGather (or the gathered pair when async_gather is False)."""
        if self.one_graph:
            self.gall.replay()
            if not gather:
                n = self.counts[self.rank]
                return self.hdr[:n], self.recs[:n]
            if self._p2p_pending is not None:
                return PendingGather(None, self._p2p_pending.bucket, list(self.counts), self.layout, self.buf)
            return PendingGather(None, self.bucket, list(self.counts), self.layout, self.buf)
        if self.split:
            return self._launch_split(async_gather, gather)
        if self._pending is not None and isinstance(self._pending, PendingGather):
            self._pending.wait()         # graph 2 rewrites self.buf: the previous volume's gather must have read it
            self._pending = None
        if self.g1 is not None:
            self.g1.replay()
        if self.world > 1:
            broadcast_prototypes(self.protos, src=self.src, group=self.eng.group)
        self.g2.replay()
        if self.world == 1 or not gather:
            n = self.counts[self.rank]
            return self.hdr[:n], self.recs[:n]
        out = gather_packed(self.buf, self.counts, self.layout, dst=self.dst, group=self.eng.group, async_op=async_gather)
        self._pending = out if async_gather else None
        return out

    def _launch_split(self, async_gather: bool, gather: bool):
        cur = torch.cuda.current_stream()
        if self._prompts_recorded:
            cur.wait_event(self._prompts_done)       # graph 2 rewrites the logits the previous prompt stage reads
        if self.g1 is not None:
            self.g1.replay()
        if self.world > 1:
            broadcast_prototypes(self.protos, src=self.src, group=self.eng.group)
        self.g2.replay()
        self._match_done.record(cur)
        out = None
        with torch.cuda.stream(self.prompt_stream):
            self.prompt_stream.wait_event(self._match_done)
            if self._pending is not None and isinstance(self._pending, PendingGather):
                self._pending.wait()                 # graph 3 rewrites self.buf: the previous gather must have read it
                self._pending = None
            self.g3.replay()
            if self.world > 1 and gather:
                out = gather_packed(self.buf, self.counts, self.layout, dst=self.dst, group=self.eng.group,
                                    async_op=async_gather)
                self._pending = out if async_gather else None
            self._prompts_done.record(self.prompt_stream)
        self._prompts_recorded = True
        if out is None:
            n = self.counts[self.rank]
            return self.hdr[:n], self.recs[:n]
        return out

    def join(self):
        """make the current stream wait for the prompt stage of the last launch() (split_streams only; a no-op otherwise)"""
        if self.split and self._prompts_recorded:
            torch.cuda.current_stream().wait_event(self._prompts_done)


# ---------------------------------------------------------------------------------------------
# Support-part management of the CT/MRI protocol
# ---------------------------------------------------------------------------------------------

def part_assign(z_id: int, z_min: int, z_max: int, npart: int) -> int:
    """Which support slice a query slice uses (dataloaders/common.py:241-249): the label's z-extent
    [z_min, z_max] is cut into `npart` equal parts; a degenerate extent maps to part 0."""
    try:
        p = int((z_id - z_min) // ((z_max - z_min) / npart))
    except ZeroDivisionError:
        p = 0
    return min(max(p, 0), npart - 1)


class PartedVolumeEngine:
    """The per-slice loop of validation_protosam.py:352-388 for one scan, batched: every label has `npart`
    support slices (one per part of its z-extent, dataloaders/common.py:228-252) and a query slice is matched,
    per label, against the support slice of the part it falls in.

    Kernel 1 runs once per part; the z-axis is then cut at every part boundary of every label, so that inside a
    segment each label has one fixed part and the whole segment goes through kernels 2-3 in one call with the
    prototype sets of those parts.  Results come back in slice order."""

    def __init__(self, engine: CoarseVolumeEngine, npart: int = 3):
        self.eng, self.npart = engine, int(npart)
        self.tables: List[dict] = []
        self.n_labels = 0

    def set_support(self, sup_feats: torch.Tensor, fg_masks: torch.Tensor):
        """sup_feats [npart,h,w,C]: the support slice of each part; fg_masks [L,npart,h,w]."""
        assert sup_feats.shape[0] == self.npart and fg_masks.shape[1] == self.npart
        self.n_labels = fg_masks.shape[0]
        self.tables = []
        for p in range(self.npart):
            self.eng.set_support(sup_feats[p:p + 1], fg_masks[:, p:p + 1].contiguous(), broadcast=False)
            self.tables.append(self.eng.protos)
        return self.tables

    def _table_for(self, parts: Sequence[int]) -> dict:
        """prototype table whose sets (bg_l, fg_l) come from part parts[l] (device-side row copies only)"""
        t0 = self.tables[0]
        L, cap, C = self.n_labels, t0["cap_rows"], t0["protos"].shape[2]
        out = ops.proto_table_alloc(2 * L, cap, C, t0["protos"].device)
        for l, p in enumerate(parts):
            src = self.tables[p]
            for k in ("protos", "counts", "eff_modes", "status"):
                out[k][2 * l: 2 * l + 2].copy_(src[k][2 * l: 2 * l + 2])
        return out

    def run(self, qry_feats: torch.Tensor, z_ids: Sequence[int], z_ranges: Sequence[Tuple[int, int]],
            return_logits: bool = False):
        """qry_feats [Q,h,w,C] with slice indices z_ids (ascending); z_ranges[l] = (z_min, z_max) of label l.
        -> (hdr [Q*L,64], recs [Q*L,max_cc,96]) on the device, image index = q*L + l, plus the part table
        [Q][L] that was used (and the coarse logits [Q*L,2,h,w] when asked for)."""
        Q, L = qry_feats.shape[0], self.n_labels
        assert len(z_ids) == Q and len(z_ranges) == L
        parts = [[part_assign(int(z), int(a), int(b), self.npart) for (a, b) in z_ranges] for z in z_ids]
        hdrs, recs, logits = [], [], []
        q0 = 0
        while q0 < Q:                                     # maximal runs of slices with identical part vectors
            q1 = q0 + 1
            while q1 < Q and parts[q1] == parts[q0]:
                q1 += 1
            self.eng.protos = self._table_for(parts[q0])
            self.eng.n_labels = L
            lg = self.eng.match(qry_feats[q0:q1])
            h, r = self.eng.prompts_from_logits(lg)
            hdrs.append(h)
            recs.append(r)
            if return_logits:
                logits.append(lg)
            q0 = q1
        out = (torch.cat(hdrs, 0), torch.cat(recs, 0), parts)
        return out + (torch.cat(logits, 0),) if return_logits else out
