"""torch-tensor front-end of the C ABI: allocates outputs/workspaces with torch (plumbing
only), passes raw device pointers + the current CUDA stream to libpsam_b200.so.

No host synchronisation happens here; every function only enqueues work on
``torch.cuda.current_stream()``.  Inputs must be CUDA float32 tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MODE_IDS

REC_DTYPE = np.dtype([("box", "<i8", 4), ("conf_pt", "<i8", 2), ("centroid", "<f8", 2), ("conf", "<f8"),
                      ("conf_pt_p", "<f4"), ("area", "<i4"), ("label", "<i4"), ("flags", "<i4"),
                      ("reserved", "<i4", 2)])
HDR_DTYPE = np.dtype([("ncc", "<i4"), ("n_rec", "<i4"), ("n_fg", "<i4"), ("flags", "<i4"),
                      ("bg_stats", "<i4", 5), ("n_runs", "<i4"), ("bg_centroid", "<f8", 2),
                      ("selected", "<i4"), ("reserved", "<i4")])
assert REC_DTYPE.itemsize == 96 and HDR_DTYPE.itemsize == 64

DEFAULT_MAX_CC = 256
DEFAULT_MAX_RUNS = 65536


def _need_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("protosam_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
        if t.dtype != torch.float32 and t.is_floating_point():
            raise RuntimeError(f"protosam_b200 computes in float32; got {t.dtype}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------ kernel 1

def proto_table_alloc(nsets, cap_rows, C, device):
    """The prototype table of psam_alp_prototypes as views of ONE byte buffer (`packed`), so that the multi-GPU
    path moves it with a single broadcast: protos [nsets,cap,C] f32 | counts | eff_modes | status [nsets] i32."""
    def up(x):
        return (x + 255) // 256 * 256
    nb_p, nb_i = up(nsets * cap_rows * C * 4), up(nsets * 4)
    packed = torch.empty(nb_p + 3 * nb_i, dtype=torch.uint8, device=device)

    def ints(k):
        return packed[nb_p + k * nb_i: nb_p + k * nb_i + nsets * 4].view(torch.int32)
    return dict(packed=packed, protos=packed[: nsets * cap_rows * C * 4].view(torch.float32).view(nsets, cap_rows, C),
                counts=ints(0), eff_modes=ints(1), status=ints(2), cap_rows=cap_rows)


def alp_prototypes(sup_x, sup_y, modes, ksize, thresh, auto_ksize=None, shots=None):
    """sup_x [S,C,h,w] (any strides; channels-last is the fast case), sup_y [nsets,S,h,w],
    modes: list of 'mask' | 'gridconv' | 'gridconv+' | 'auto_fg'; shots: optional list, per set -1 (every shot)
    or the one shot the set is restricted to.  Returns a dict of device tensors (see include/psam_b200.h,
    psam_alp_prototypes / psam_alp_prototypes_shots)."""
    L = _lib.load()
    _need_cuda(sup_x, sup_y)
    S, C, h, w = sup_x.shape
    sup_y = sup_y.reshape(-1, S, h, w).contiguous()
    nsets = sup_y.shape[0]
    assert len(modes) == nsets
    kh, kw = int(ksize[0]), int(ksize[1])
    akh, akw = (int(auto_ksize[0]), int(auto_ksize[1])) if auto_ksize is not None else (kh, kw)
    gh, gw = h // kh, w // kw
    N = S * gh * gw
    cap = N + S
    dev = sup_x.device
    out = proto_table_alloc(nsets, cap, C, dev)
    out.update(
        survive=torch.empty((nsets, max(N, 1)), dtype=torch.uint8, device=dev),
        pooled=torch.empty((nsets, max(N, 1)), dtype=torch.float32, device=dev),
        N=N, gh=gh, gw=gw, S=S, C=C,
    )
    ws = _ws(L.psam_alp_prototypes_workspace(nsets, S, C, h, w, kh, kw), dev)
    strides = (ctypes.c_int64 * 4)(*sup_x.stride())
    mode_arr = (ctypes.c_int32 * nsets)(*[MODE_IDS[m] if isinstance(m, str) else int(m) for m in modes])
    tail = (S, C, h, w, kh, kw, akh, akw, float(thresh), _ptr(out["protos"]), _ptr(out["counts"]),
            _ptr(out["eff_modes"]), _ptr(out["status"]), _ptr(out["survive"]), _ptr(out["pooled"]), _ptr(ws), ws.numel(),
            _stream())
    if shots is None:
        rc = L.psam_alp_prototypes(_ptr(sup_x), strides, _ptr(sup_y), nsets, mode_arr, *tail)
    else:
        assert len(shots) == nsets
        shot_arr = (ctypes.c_int32 * nsets)(*[int(v) for v in shots])
        rc = L.psam_alp_prototypes_shots(_ptr(sup_x), strides, _ptr(sup_y), nsets, mode_arr, shot_arr, *tail)
    _lib.check(rc, "psam_alp_prototypes")
    out["_ws"] = ws          # keep alive until the stream has consumed it
    return out


def combine_shots(scores, L, S):
    """scores [Q, L*(1+S), HW] with sets (bg_l, fg_l shot 0..S-1) -> logits [Q*L, 2, HW] = (bg_l, max over shots of
    fg_l), the element-wise max of grid_proto_fewshot.py:263-270."""
    lib = _lib.load()
    _need_cuda(scores)
    scores = scores.contiguous()
    Q, nsets, HW = scores.shape
    assert nsets == L * (1 + S)
    out = torch.empty((Q * L, 2, HW), dtype=torch.float32, device=scores.device)
    rc = lib.psam_combine_shots(_ptr(scores), Q, L, S, HW, _ptr(out), _stream())
    _lib.check(rc, "psam_combine_shots")
    return out


DEFAULT_FEATURE_SIZE = 32          # util/consts.py:1


def tokens_to_features(tokens, feature_size=DEFAULT_FEATURE_SIZE):
    """FewShotSeg.get_features' hand-off (grid_proto_fewshot.py:90-98): x_norm_patchtokens [B, HW, C] -> channels-last
    features [B, h, w, C]; maps with fewer than feature_size^2 tokens are resized bilinearly to feature_size^2."""
    L = _lib.load()
    _need_cuda(tokens)
    B, HW, C = tokens.shape
    h = w = int(HW ** 0.5)
    tokens = tokens.contiguous()
    if HW >= feature_size ** 2:
        return tokens.view(B, h, w, C)
    out = torch.empty((B, feature_size, feature_size, C), dtype=torch.float32, device=tokens.device)
    rc = L.psam_tokens_to_features(_ptr(tokens), B, h, w, C, feature_size, feature_size, _ptr(out), _stream())
    _lib.check(rc, "psam_tokens_to_features")
    return out


def mask_nearest(masks, h, w):
    """F.interpolate(masks, (h, w), mode='nearest') for float masks [..., H, W] (grid_proto_fewshot.py:228-231)."""
    L = _lib.load()
    _need_cuda(masks)
    src = masks.contiguous()
    H, W = src.shape[-2:]
    n = src.numel() // (H * W)
    dst = torch.empty(tuple(src.shape[:-2]) + (h, w), dtype=torch.float32, device=src.device)
    rc = L.psam_mask_nearest(_ptr(src), n, H, W, int(h), int(w), _ptr(dst), _stream())
    _lib.check(rc, "psam_mask_nearest")
    return dst


def alp_proto_grid(pooled_set, S, gh, gw, vw, thresh, mode):
    """`resized_proto_grid` [1,1,gh*vw,gw*vw] of one set (viz only)."""
    L = _lib.load()
    _need_cuda(pooled_set)
    out = torch.zeros((1, 1, gh * vw, gw * vw), dtype=torch.float32, device=pooled_set.device)
    rc = L.psam_alp_proto_grid(_ptr(pooled_set), S, gh, gw, vw, float(thresh), MODE_IDS[mode], _ptr(out), _stream())
    _lib.check(rc, "psam_alp_proto_grid")
    return out


# ------------------------------------------------------------------ kernel 2

def alp_match(qry, protos, want_assign=True, want_sims=False, algo=0):
    """qry [Q,HW,C] with unit channel stride (rows may be padded / slices strided);
    protos = dict from alp_prototypes.  Returns (scores [Q,nsets,HW], assign | None, sims | None)."""
    L = _lib.load()
    _need_cuda(qry)
    Q, HW, C = qry.shape
    if qry.stride(2) != 1:
        raise RuntimeError("alp_match needs channels-last rows (stride of C == 1)")
    nsets, cap = protos["protos"].shape[0], protos["cap_rows"]
    dev = qry.device
    scores = torch.empty((Q, nsets, HW), dtype=torch.float32, device=dev)
    assign = torch.empty((Q, nsets, HW), dtype=torch.float32, device=dev) if want_assign else None
    sims = torch.empty((Q, nsets, cap, HW), dtype=torch.float32, device=dev) if want_sims else None
    ws = _ws(L.psam_alp_match_workspace(Q, HW, C, nsets, cap, algo), dev)
    rc = L.psam_alp_match(_ptr(qry), qry.stride(0), qry.stride(1), Q, HW, C, _ptr(protos["protos"]), cap,
                          _ptr(protos["counts"]), _ptr(protos["eff_modes"]), nsets, _ptr(scores), _ptr(assign),
                          _ptr(sims), _ptr(protos["status"]), _ptr(ws), ws.numel(), algo, _stream())
    _lib.check(rc, "psam_alp_match")
    return scores, assign, sims


# ------------------------------------------------------------------ kernel 3

PROB_MODES = {"softmax": _lib.PROB_SOFTMAX, "softmax_twice": _lib.PROB_SOFTMAX_TWICE}


def upsample_softmax(logits, mid, out=1024, want_p_fg=True, want_probs2=False, fg_only=False, want_wstat=False,
                     prob_mode="softmax"):
    """logits [n,2,h,w] -> (p_fg [n,out,out] | None, maskbits [n,out,ceil(out/32)] int32, probs2 | None
    [, wstat int64 [n,out,ceil(out/32)]]).  fg_only: the engine's variant (p_fg written only where kernel 3b reads it).
    prob_mode 'softmax_twice': p_fg / wstat carry softmax(softmax(logits))[1], ProtoMedSAM's confidence map."""
    L = _lib.load()
    _need_cuda(logits)
    logits = logits.contiguous()
    n, two, h, w = logits.shape
    assert two == 2
    dev = logits.device
    p_fg = torch.empty((n, out, out), dtype=torch.float32, device=dev) if want_p_fg else None
    wpr = (out + 31) // 32
    bits = torch.empty((n, out, wpr), dtype=torch.int32, device=dev)
    probs2 = torch.empty((n, 2, out, out), dtype=torch.float32, device=dev) if want_probs2 else None
    wstat = torch.zeros((n, out, wpr), dtype=torch.int64, device=dev) if want_wstat else None
    if fg_only and p_fg is not None:
        p_fg.zero_()
    ws = _ws(L.psam_upsample_workspace(n, int(out)), dev)
    rc = L.psam_upsample_softmax(_ptr(logits), n, h, w, int(mid), int(out), _ptr(p_fg), _ptr(bits), _ptr(probs2),
                                 _ptr(wstat), int(bool(fg_only)), PROB_MODES[prob_mode], _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "psam_upsample_softmax")
    if want_wstat:
        return p_fg, bits, probs2, wstat
    return p_fg, bits, probs2


def components(maskbits, p_fg, use_cca=False, max_cc=DEFAULT_MAX_CC, max_runs=DEFAULT_MAX_RUNS, want_labels=False,
               wstat=None):
    """-> (hdr uint8 [n,64], recs uint8 [n,max_cc,96], labels int32 [n,out,out] | None), all on device."""
    L = _lib.load()
    _need_cuda(p_fg)
    n, out, _ = p_fg.shape
    dev = p_fg.device
    hdr = torch.zeros((n, HDR_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    recs = torch.zeros((n, max_cc, REC_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    labels = torch.empty((n, out, out), dtype=torch.int32, device=dev) if want_labels else None
    ws = _ws(L.psam_prompts_workspace(n, out, max_runs, max_cc), dev)
    rc = L.psam_components(_ptr(maskbits), _ptr(p_fg), _ptr(wstat), n, out, int(bool(use_cca)), max_cc, max_runs, _ptr(hdr),
                           _ptr(recs), _ptr(labels), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "psam_components")
    return hdr, recs, labels


def records_alloc(n_alloc, max_cc, device):
    """Headers [n,64] and records [n,max_cc,96] as views of ONE zeroed byte buffer (one gather moves both)."""
    buf = torch.zeros(n_alloc * (HDR_DTYPE.itemsize + max_cc * REC_DTYPE.itemsize), dtype=torch.uint8, device=device)
    hdr = buf[: n_alloc * HDR_DTYPE.itemsize].view(n_alloc, HDR_DTYPE.itemsize)
    recs = buf[n_alloc * HDR_DTYPE.itemsize:].view(n_alloc, max_cc, REC_DTYPE.itemsize)
    return buf, hdr, recs


def split_records(buf, max_cc):
    """inverse of records_alloc's layout for a received buffer"""
    n_alloc = buf.numel() // (HDR_DTYPE.itemsize + max_cc * REC_DTYPE.itemsize)
    return (buf[: n_alloc * HDR_DTYPE.itemsize].view(n_alloc, HDR_DTYPE.itemsize),
            buf[n_alloc * HDR_DTYPE.itemsize:].view(n_alloc, max_cc, REC_DTYPE.itemsize))


TAIL_DTYPE = np.dtype([("total", "<i4"), ("capacity", "<i4"), ("flags", "<i4"), ("n_img", "<i4"), ("reserved", "<i4", 12)])
assert TAIL_DTYPE.itemsize == 64
DEFAULT_RECS_PER_IMAGE = 6        # capacity of the compact form: live records per image on average (typical: 1-3)


def packed_bytes(n_alloc, capacity):
    return n_alloc * HDR_DTYPE.itemsize + TAIL_DTYPE.itemsize + capacity * REC_DTYPE.itemsize


def compact_records(hdr, recs, n_alloc=None, capacity=None, out=None):
    """Dense (hdr [n,64], recs [n,max_cc,96]) -> ONE compact byte buffer [n_alloc headers | tail | capacity records]
    (psam_compact_records): what a gather or a device->host copy moves.  hdr.reserved = index of the image's first record."""
    L = _lib.load()
    n, max_cc = recs.shape[0], recs.shape[1]
    n_alloc = max(int(n_alloc or n), n)
    capacity = int(capacity) if capacity else n_alloc * DEFAULT_RECS_PER_IMAGE
    nb = packed_bytes(n_alloc, capacity)
    if out is None:
        out = torch.empty(nb, dtype=torch.uint8, device=recs.device)
    assert out.numel() == nb
    rc = L.psam_compact_records(_ptr(hdr), _ptr(recs), n, n_alloc, max_cc, capacity, _ptr(out), _stream())
    _lib.check(rc, "psam_compact_records")
    return out


def split_packed(buf, n_alloc, capacity):
    """views of a compact buffer: (hdr [n_alloc,64], tail [64], recs [capacity,96])"""
    a = n_alloc * HDR_DTYPE.itemsize
    return (buf[:a].view(n_alloc, HDR_DTYPE.itemsize), buf[a: a + TAIL_DTYPE.itemsize],
            buf[a + TAIL_DTYPE.itemsize:].view(capacity, REC_DTYPE.itemsize))


def decode_packed(buf, n_alloc, capacity, n_valid=None):
    """compact buffer (device or host) -> (headers [n_valid], records [total]) as numpy structured arrays; image i owns
    records[H['reserved'][i] : H['reserved'][i] + H['n_rec'][i]]."""
    raw = buf.detach().cpu().numpy().tobytes()
    a = n_alloc * HDR_DTYPE.itemsize
    H = np.frombuffer(raw[:a], dtype=HDR_DTYPE)
    T = np.frombuffer(raw[a: a + TAIL_DTYPE.itemsize], dtype=TAIL_DTYPE)[0]
    if int(T["flags"]) & _lib.PACKED_OVERFLOW:
        raise RuntimeError(f"{int(T['total'])} prompt records exceed the compact buffer's capacity {int(T['capacity'])}; "
                           "raise recs_per_image (or read the dense records)")
    R = np.frombuffer(raw[a + TAIL_DTYPE.itemsize:], dtype=REC_DTYPE)[: int(T["total"])]
    return H[: (n_valid if n_valid is not None else int(T["n_img"]))], R


def coarse_to_prompts(logits, mid, out=1024, use_cca=False, max_cc=DEFAULT_MAX_CC, max_runs=DEFAULT_MAX_RUNS,
                      workspace=None, n_alloc=None, return_packed=False, prob_mode="softmax", capacity=None):
    """logits [n,2,h,w] -> (hdr uint8 [n,64], recs uint8 [n,max_cc,96]) on device, one call.  return_packed adds the
    compact buffer (compact_records) sized for n_alloc >= n images, e.g. the largest shard of a multi-GPU run."""
    L = _lib.load()
    _need_cuda(logits)
    logits = logits.contiguous()
    n, two, h, w = logits.shape
    assert two == 2
    dev = logits.device
    _, hdr, recs = records_alloc(n, max_cc, dev)
    need = L.psam_coarse_to_prompts_workspace(n, out, max_runs, max_cc)
    ws = workspace if workspace is not None and workspace.numel() >= need else _ws(need, dev)
    rc = L.psam_coarse_to_prompts(_ptr(logits), n, h, w, int(mid), int(out), int(bool(use_cca)), PROB_MODES[prob_mode],
                                  max_cc, max_runs, _ptr(hdr), _ptr(recs), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "psam_coarse_to_prompts")
    if return_packed:
        return hdr, recs, compact_records(hdr, recs, n_alloc=n_alloc, capacity=capacity)
    return hdr, recs


NEG_DTYPE = np.dtype([("pt", "<i8", 2), ("p", "<f4"), ("has", "<i4"), ("n", "<i4"), ("reserved", "<i4")])
assert NEG_DTYPE.itemsize == 32


def neg_points(labels, p_bg, hdr, recs, use_cca=False, ring_width=10, thresh=0.95, host_aliasing=False):
    """Negative-point candidates (ProtoSAM.get_sam_input_points with get_neg_points=True, models/ProtoSAM.py:361-434):
    labels int32 [n,out,out] (psam_components), p_bg [n,out,out] view of the background probabilities (its image
    stride may be larger: channel 0 of probs2) -> uint8 [n, max_cc + 1, 32] records (NEG_DTYPE): [i, r] the ring point
    of component r, [i, max_cc] the image's global point."""
    L = _lib.load()
    _need_cuda(p_bg)
    n, out, _ = labels.shape
    max_cc = recs.shape[1]
    assert p_bg.stride(2) == 1 and p_bg.stride(1) == out and labels.is_contiguous()
    neg = torch.empty((n, max_cc + 1, NEG_DTYPE.itemsize), dtype=torch.uint8, device=labels.device)
    rc = L.psam_neg_points(_ptr(labels), _ptr(p_bg), p_bg.stride(0), _ptr(hdr.contiguous()), _ptr(recs.contiguous()), n, out,
                           max_cc, int(bool(use_cca)), int(ring_width), float(thresh), int(bool(host_aliasing)), _ptr(neg),
                           _stream())
    _lib.check(rc, "psam_neg_points")
    return neg


# ---------------------------------------------------------------------------------------------
# One-sided exchanges over NVLink peer memory (include/psam_b200.h, psam_peer_*)
# ---------------------------------------------------------------------------------------------

class PeerChannel:
    """One symmetric region per rank ([signal words | mailbox of payload_bytes]) mapped into every peer through
    torch.distributed._symmetric_memory, plus the local epoch counters of the two directions.  Creating one is a
    collective, blocking call (allocation + rendezvous + barrier): do it outside CUDA-graph capture."""

    def __init__(self, payload_bytes: int, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        L = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.payload_bytes = int(payload_bytes)
        nbytes = int(L.psam_peer_region_bytes(self.payload_bytes))
        self.region = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.region.zero_()
        self.handle = symm_mem.rendezvous(self.region, self.group)
        self.regions_dev = int(self.handle.buffer_ptrs_dev)          # device array of the world's region addresses
        self.ctr = torch.zeros(8, dtype=torch.int32, device=device)  # [send epoch, done, pad, pad | recv epoch, done, pad, pad]
        torch.cuda.current_stream().synchronize()
        self.handle.barrier()                                        # nobody signals into a region that is not zeroed yet
        torch.cuda.current_stream().synchronize()

    def _ctr(self, recv: bool) -> int:
        return self.ctr.data_ptr() + (16 if recv else 0)


def _table_geom(protos):
    packed, rows = protos["packed"], protos["protos"]
    nsets, cap, C = rows.shape
    ints_offset = protos["counts"].data_ptr() - packed.data_ptr()
    return nsets, cap, C, ints_offset, packed.numel() - ints_offset


def peer_push_table(ch: PeerChannel, protos):
    """source rank: live prototype rows + integer arrays -> every peer's mailbox (waits for their previous acknowledgement)"""
    L = _lib.load()
    nsets, cap, C, io, ib = _table_geom(protos)
    assert protos["packed"].numel() <= ch.payload_bytes
    rc = L.psam_peer_push_table(_ptr(protos["packed"]), nsets, cap, C, io, ib, ch.regions_dev, ch.world, ch.rank, ch._ctr(False),
                                _stream())
    _lib.check(rc, "psam_peer_push_table")


def peer_recv_table(ch: PeerChannel, protos, src: int):
    """other ranks: wait for the source's table, copy it out of the mailbox into `protos`, acknowledge"""
    L = _lib.load()
    nsets, cap, C, io, ib = _table_geom(protos)
    assert protos["packed"].numel() <= ch.payload_bytes
    rc = L.psam_peer_recv_table(_ptr(protos["packed"]), nsets, cap, C, io, ib, ch.regions_dev, ch.world, ch.rank, int(src),
                                ch._ctr(False), _stream())
    _lib.check(rc, "psam_peer_recv_table")


def peer_put(ch: PeerChannel, buf, dst: int):
    """every rank: `buf` (uint8, a multiple of 16 bytes) -> slot `rank` of dst's mailbox"""
    L = _lib.load()
    slot = ch.payload_bytes // ch.world // 16 * 16
    assert buf.is_contiguous() and buf.numel() % 16 == 0 and buf.numel() <= slot
    rc = L.psam_peer_put(_ptr(buf), buf.numel(), slot, ch.regions_dev, ch.world, ch.rank, int(dst), ch._ctr(False), _stream())
    _lib.check(rc, "psam_peer_put")


def peer_collect(ch: PeerChannel, out):
    """dst: wait for every rank's slot, copy the mailbox -> out (uint8 [world, slot_bytes]), acknowledge to all"""
    L = _lib.load()
    slot = ch.payload_bytes // ch.world // 16 * 16
    assert out.is_contiguous() and out.shape == (ch.world, slot)
    rc = L.psam_peer_collect(_ptr(out), slot, ch.regions_dev, ch.world, ch.rank, ch._ctr(False), ch._ctr(True), _stream())
    _lib.check(rc, "psam_peer_collect")


def topk_points(labels, p_fg, hdr, recs, k, use_cca=False):
    """get_most_conf_points for any k (models/ProtoSAM.py:266-289): -> (pts int64 [n,max_cc,k,2] in (x, y), conf float32
    [n,max_cc,k]), torch.topk's order within each component; -1 where a component has fewer than k pixels."""
    L = _lib.load()
    _need_cuda(p_fg)
    n, out, _ = labels.shape
    max_cc = recs.shape[1]
    assert p_fg.stride(2) == 1 and p_fg.stride(1) == out and labels.is_contiguous()
    pts = torch.empty((n, max_cc, k, 2), dtype=torch.int64, device=labels.device)
    conf = torch.empty((n, max_cc, k), dtype=torch.float32, device=labels.device)
    ws = _ws(L.psam_topk_points_workspace(n, max_cc, int(k)), labels.device)
    rc = L.psam_topk_points(_ptr(labels), _ptr(p_fg), p_fg.stride(0), _ptr(hdr.contiguous()), _ptr(recs.contiguous()), n, out,
                            max_cc, int(bool(use_cca)), int(k), _ptr(pts), _ptr(conf), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "psam_topk_points")
    return pts, conf


def mask_prompts(labels, hdr, recs, use_cca=False, size=256, capacity=None):
    """Mask prompts (models/ProtoSAM.py:452-476): -> (masks uint8 [capacity,size,size] with 10 / 248, offsets int32 [n+1]);
    component r of image i is masks[offsets[i] + r]."""
    L = _lib.load()
    n, out, _ = labels.shape
    max_cc = recs.shape[1]
    capacity = int(capacity) if capacity is not None else n * max_cc
    masks = torch.empty((capacity, size, size), dtype=torch.uint8, device=labels.device)
    offsets = torch.zeros(n + 1, dtype=torch.int32, device=labels.device)
    rc = L.psam_mask_prompts(_ptr(labels), _ptr(hdr.contiguous()), _ptr(recs.contiguous()), n, out, max_cc,
                             int(bool(use_cca)), int(size), capacity, _ptr(masks), _ptr(offsets), _stream())
    _lib.check(rc, "psam_mask_prompts")
    return masks, offsets


def confidence(p_fg):
    """get_confidence_from_logits (util/utils.py:429-434) from p_fg [n, ...]: float64 [n] on the device."""
    L = _lib.load()
    _need_cuda(p_fg)
    p_fg = p_fg.contiguous()
    n = p_fg.shape[0]
    conf = torch.empty(n, dtype=torch.float64, device=p_fg.device)
    rc = L.psam_confidence(_ptr(p_fg), n, p_fg.numel() // n, _ptr(conf), _stream())
    _lib.check(rc, "psam_confidence")
    return conf


POINT_MODE_IDS = {"conf": 0, "centroid": 1, "both": 2}


def records_to_sam(hdr, recs, point_mode="both", original_size=(1024, 1024), target_length=1024):
    """Device records -> (points [n,max_cc,npts,2] f32, labels [n,max_cc,npts] i32, boxes [n,max_cc,4] f32) in SAM's
    input frame, ready for SamPredictor.predict_torch (one call per image with batch = its n_rec components).  No host
    round trip: the prompts never leave the device.  Slots beyond an image's n_rec are zero; an image whose header
    carries IMG_CC_TRUNCATED (more components than max_cc) or IMG_RUN_OVERFLOW has FEWER prompts here than the reference
    would issue -- the flags stay in hdr[:, 12:16] (int32 `flags`) for the caller to test on the device, and
    CoarseVolumeEngine.decode() raises on them."""
    L = _lib.load()
    hdr, recs = hdr.contiguous(), recs.contiguous()
    n, max_cc = recs.shape[0], recs.shape[1]
    npts = 2 if point_mode == "both" else 1
    dev = recs.device
    points = torch.empty((n, max_cc, npts, 2), dtype=torch.float32, device=dev)
    labels = torch.empty((n, max_cc, npts), dtype=torch.int32, device=dev)
    boxes = torch.empty((n, max_cc, 4), dtype=torch.float32, device=dev)
    rc = L.psam_records_to_sam(_ptr(hdr), _ptr(recs), n, max_cc, POINT_MODE_IDS[point_mode], int(original_size[0]),
                               int(original_size[1]), int(target_length), _ptr(points), _ptr(labels), _ptr(boxes), _stream())
    _lib.check(rc, "psam_records_to_sam")
    return points, labels, boxes


def decode_headers(hdr_u8) -> np.ndarray:
    """device/host uint8 [n,64] -> numpy structured array (host sync on .cpu())."""
    return np.frombuffer(hdr_u8.detach().cpu().numpy().tobytes(), dtype=HDR_DTYPE)


def decode_records(recs_u8) -> np.ndarray:
    a = recs_u8.detach().cpu().numpy()
    return np.frombuffer(a.tobytes(), dtype=REC_DTYPE).reshape(a.shape[0], a.shape[1])
