"""numpy/ctypes front-end of the CPU oracle (oracle/psam_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of psam_oracle.c.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs import this module.  ``protosam_b200`` never does.

Every function names the reference lines it restates (paths relative to the
levayz/ProtoSAM tree).  Parity is pinned by execution: the fixtures under
``tests/golden/`` are outputs of the unmodified reference (see
``oracle/make_golden.py``) and ``tests/test_oracle_golden.py`` checks this
module against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpsam_oracle.so")

MODE_IDS = {"mask": 0, "gridconv": 1, "gridconv+": 2}

_P = ctypes.c_void_p
_lib = None


def build(force: bool = False) -> str:
    """Compile psam_oracle.c with gcc (a few hundred ms).  Building the checker
    is not using it."""
    src = os.path.join(_HERE, "psam_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.psamo_prototypes.restype = ctypes.c_int
        L.psamo_match.restype = ctypes.c_int
        L.psamo_ccl8.restype = ctypes.c_int
        L.psamo_pairwise_sum_f32.restype = ctypes.c_float
        L.psamo_expf_u10.restype = ctypes.c_float
        L.psamo_expf_u10.argtypes = [ctypes.c_float]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(_P)


def _elem_strides(a, dims):
    return np.array([a.strides[d] // a.itemsize for d in dims], dtype=np.int64)


# --------------------------------------------------------------------------
# A. prototypes -- models/alpmodule.py:97-159
# --------------------------------------------------------------------------

def get_prototypes(sup_x, sup_y, mode, ksize, thresh, val_wsize=None):
    """``MultiProtoAsConv.get_prototypes`` (models/alpmodule.py:97-159).

    sup_x [S,C,h,w] float32 (any strides), sup_y [S,1,h,w]; ksize = (kh, kw) is
    the pooling window actually used (``val_wsize`` twice when ``isval`` else
    ``self.kernel_size``); val_wsize = the scalar the reference uses to size
    ``resized_proto_grid`` (:126,148; defaults to kh).

    Returns dict(pro_n [P,C], survive [S*gh*gw] bool, pooled [S,1,gh,gw],
    proto_grid (the returned ``resized_proto_grid``), non_zero [nnz,4] int64).
    """
    L = lib()
    sup_x = np.asarray(sup_x, dtype=np.float32)
    S, C, h, w = sup_x.shape
    y = _f32(np.asarray(sup_y).reshape(S, h, w))
    kh, kw = int(ksize[0]), int(ksize[1])
    vw = kh if val_wsize is None else int(val_wsize)
    m = MODE_IDS[mode]
    gh, gw = h // kh, w // kw
    N = S * gh * gw
    protos = np.zeros((N + S, C), np.float32)
    survive = np.zeros(max(N, 1), np.uint8)
    pooled = np.zeros(max(N, 1), np.float32)
    xs = _elem_strides(sup_x, (0, 1, 2, 3))
    rows = L.psamo_prototypes(_ptr(sup_x), _ptr(xs), _ptr(y), S, C, h, w, kh, kw,
                              ctypes.c_float(thresh), m, _ptr(protos), _ptr(survive), _ptr(pooled))
    out = {"pro_n": protos[:rows].copy()}
    if mode == "mask":
        # :104-106  proto_grid = sup_y.clone(); non_zero = nonzero(proto_grid)
        grid = y.reshape(S, 1, h, w).copy()
        out.update(survive=np.zeros(0, bool), pooled=np.zeros((S, 1, 0, 0), np.float32),
                   proto_grid=grid, non_zero=np.argwhere(grid != 0).astype(np.int64))
        return out
    pooled = pooled[:N].reshape(S, 1, gh, gw)
    out["survive"] = survive[:N].astype(bool)
    out["pooled"] = pooled
    # viz grid, :120-128 / :142-150
    grid = pooled.copy()
    grid[grid < np.float32(thresh)] = 0
    non_zero = np.argwhere(grid != 0).astype(np.int64)
    if mode == "gridconv+":
        for i, idx in enumerate(non_zero):          # :146-147 (first index is 0, not idx[0])
            grid[0, idx[1], idx[2], idx[3]] = i + 1
    resized = np.zeros((1, 1, gh * vw, gw * vw), np.float32)
    for idx in non_zero:                            # :127-128 / :149-150 (column extent hard-coded 2)
        resized[0, 0, idx[2] * vw:idx[2] * vw + vw, idx[3] * vw:idx[3] * vw + 2] = grid[0, 0, idx[2], idx[3]]
    out["proto_grid"] = resized
    out["non_zero"] = non_zero
    return out


# --------------------------------------------------------------------------
# B. match -- models/alpmodule.py:57-94
# --------------------------------------------------------------------------

def get_prediction(pro_n, qry, mode, want_sims=False):
    """``get_prediction_from_prototypes`` (models/alpmodule.py:57-94) with the
    query normalisation of ``forward`` (:195) folded in.

    qry [C,h,w] (any strides); pro_n [P,C].  Returns (pred [h,w], assign [h,w],
    sims [P,h,w] | [h,w] | None).  Raises RuntimeError when P == 0 like the
    reference's F.conv2d does."""
    L = lib()
    qry = np.asarray(qry, dtype=np.float32)
    C, h, w = qry.shape
    pro_n = _f32(pro_n)
    P = pro_n.shape[0]
    if P == 0:
        raise RuntimeError("no prototypes: conv2d with a [0,C,1,1] weight (models/alpmodule.py:68)")
    m = MODE_IDS[mode]
    pred = np.empty((h, w), np.float32)
    assign = np.empty((h, w), np.float32)
    sims = None
    if want_sims:
        sims = np.empty((h, w) if mode == "mask" else (P, h, w), np.float32)
    qs = _elem_strides(qry, (0, 1, 2))
    rc = L.psamo_match(_ptr(qry), _ptr(qs), C, h, w, _ptr(pro_n), P, m, _ptr(pred), _ptr(assign),
                       _ptr(sims) if sims is not None else None)
    assert rc == 0
    return pred, assign, sims


def alp_forward(qry, sup_x, sup_y, mode, thresh, kernel_size, isval=False, val_wsize=None, vis_sim=False):
    """``MultiProtoAsConv.forward`` (models/alpmodule.py:161-198) on numpy arrays.

    qry [1,C,h,w] or [1,1,C,h,w]; sup_x [1,S,1,C,h,w]; sup_y [1,S,1,h,w].
    Returns (pred_grid [1,1,h,w], [debug_assign [1,h,w]], vis_dict, proto_grid)."""
    if mode not in MODE_IDS:
        raise ValueError(f"Invalid mode: {mode}. Expected 'mask', 'gridconv', or 'gridconv+'.")
    qry = np.asarray(qry)
    if qry.ndim == 5:
        qry = qry[:, 0]
    sup_x = np.asarray(sup_x)
    sup_x = sup_x.reshape(sup_x.shape[1], *sup_x.shape[3:])            # :179
    S, C, h, w = sup_x.shape
    if val_wsize is None:                                              # :187-190
        val_wsize = kernel_size[0]
        ks = (kernel_size[0], kernel_size[1])
    else:
        ks = (val_wsize, val_wsize) if isval else (kernel_size[0], kernel_size[1])
    protos = get_prototypes(sup_x, np.asarray(sup_y).reshape(S, 1, h, w), mode, ks, thresh, val_wsize)
    pred, assign, sims = get_prediction(protos["pro_n"], qry[0], mode, want_sims=vis_sim)
    vis = {"proto_assign": assign[None]}
    if vis_sim:
        vis["raw_local_sims"] = sims[None]
    return pred[None, None], [assign[None]], vis, protos["proto_grid"]


# --------------------------------------------------------------------------
# C. coarse map -> probabilities -- grid_proto_fewshot.py:270-273, ProtoSAM.py:592-602
# --------------------------------------------------------------------------

def upsample_bilinear(x, size):
    """``F.interpolate(x, size, mode='bilinear')`` for [N,C,ih,iw] float32, ATen CPU
    arithmetic (see psamo_bilinear)."""
    L = lib()
    x = _f32(x)
    N, C, ih, iw = x.shape
    oh, ow = (size, size) if np.isscalar(size) else size
    # the path only ever upsamples (feature grid -> image -> 1024); ATen's arithmetic
    # for out < in was not pinned and is rejected rather than guessed
    assert oh >= ih and ow >= iw, "oracle bilinear restates upsampling only"
    out = np.empty((N, C, oh, ow), np.float32)
    for n in range(N):
        for c in range(C):
            L.psamo_bilinear(_ptr(x[n, c]), ih, iw, _ptr(out[n, c]), oh, ow)
    return out


def softmax2(logits):
    """``logits.softmax(dim=1)`` for [1,2,H,W] float32 (ProtoSAM.py:599)."""
    L = lib()
    logits = _f32(logits)
    assert logits.shape[0] == 1 and logits.shape[1] == 2
    out = np.empty_like(logits)
    n = logits.shape[2] * logits.shape[3]
    L.psamo_softmax2(_ptr(logits[0, 0]), _ptr(logits[0, 1]), _ptr(out[0, 0]), _ptr(out[0, 1]), ctypes.c_int64(n))
    return out


def coarse_logits_to_probs(low_logits, mid_size, out_size=1024):
    """[1,2,h,w] raw scores -> (output_logits [1,2,out,out], output_p, pred uint8 [out,out]).

    Stage 1: FewShotSeg upsamples to the ALPNet image size (grid_proto_fewshot.py:270-273);
    stage 2: ProtoSAM upsamples to its own image_size when it differs (ProtoSAM.py:592-594);
    then softmax + argmax (:599-602; argmax over two classes keeps class 0 on ties)."""
    x = upsample_bilinear(low_logits, mid_size)
    if mid_size != out_size:
        x = upsample_bilinear(x, out_size)
    p = softmax2(x)
    pred = (p[0, 1] > p[0, 0]).astype(np.uint8)
    return x, p, pred


# --------------------------------------------------------------------------
# D. connected components + confidences -- util/utils.py:474-541
# --------------------------------------------------------------------------

def connected_components(mask):
    """``cv2.connectedComponentsWithStats(mask.astype(np.uint8), connectivity=8)``
    (util/utils.py:478): (n, labels int32, stats int32 [n,5], centroids float64 [n,2])."""
    L = lib()
    mask = np.ascontiguousarray(mask).astype(np.uint8)
    H, W = mask.shape
    labels = np.empty((H, W), np.int32)
    cap = 1024
    while True:
        stats = np.zeros((cap, 5), np.int32)
        cent = np.zeros((cap, 2), np.float64)
        n = L.psamo_ccl8(_ptr(mask), H, W, _ptr(labels), _ptr(stats), _ptr(cent), cap)
        if n >= 0:
            break
        cap = -n
    return n, labels, stats[:n].copy(), cent[:n].copy()


def cc_sums(p_fg, labels, nlab):
    """float32 ``(p_fg.flatten() * (labels == j).flatten()).sum()`` for every j
    (numpy pairwise order, util/utils.py:490)."""
    L = lib()
    p_fg = _f32(p_fg)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    sums = np.zeros(nlab, np.float32)
    L.psamo_cc_sums(_ptr(p_fg), _ptr(labels), ctypes.c_int64(p_fg.size), nlab, _ptr(sums))
    return sums


def get_connected_components(pred, p_fg, return_conf=False):
    """util/utils.py:474-494.  ``p_fg`` is ``query_pred_logits.softmax(1)[:,1]``
    already evaluated (the caller holds it)."""
    cc = connected_components(pred)
    if not return_conf:
        return cc, None
    sums = cc_sums(p_fg, cc[1], cc[0])
    denom = np.float64(np.asarray(pred).astype(np.int64).sum()) + 1e-6
    conf = {0: 0}
    for j in range(1, cc[0]):
        conf[j] = np.float64(sums[j]) / denom
    return cc, conf


def cca(pred, p_fg, return_conf=False, return_cc=False):
    """util/utils.py:496-541: keep the component with the strictly largest confidence."""
    pred = np.asarray(pred)
    cc, conf = get_connected_components(pred, p_fg, return_conf=True)
    max_conf, max_key = conf[0], None
    for k, v in conf.items():
        if v > max_conf:
            max_conf, max_key = v, k
    if max_conf == 0:
        query_pred = np.zeros_like(pred)
    else:
        cc = (2, np.where(cc[1] != max_key, 0, 1), cc[2][[0, max_key]], cc[3][[0, max_key]])
        query_pred = (cc[1] == 1).astype(np.uint8)
    if return_cc:
        return cc
    out = pred * query_pred
    if return_conf:
        return out, max_conf
    return out


# --------------------------------------------------------------------------
# E. prompts -- models/ProtoSAM.py:242-289, 349-450
# --------------------------------------------------------------------------

def _cc_prompts(p_fg, labels, nlab):
    L = lib()
    p_fg = _f32(p_fg)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    H, W = labels.shape
    bbox = np.zeros((nlab, 4), np.int64)
    pt = np.zeros((nlab, 2), np.int64)
    val = np.zeros(nlab, np.float32)
    L.psamo_cc_prompts(_ptr(p_fg), _ptr(labels), H, W, nlab, _ptr(bbox), _ptr(pt), _ptr(val))
    return bbox, pt, val


def get_bbox_per_cc(cc):
    """models/ProtoSAM.py:242-264 -> int64 [n-1,4] XYXY inclusive."""
    bbox, _, _ = _cc_prompts(np.zeros(cc[1].shape, np.float32), cc[1], cc[0])
    return bbox[1:]


def get_sam_input_points(cc, p_fg, point_mode="both", k=1):
    """models/ProtoSAM.py:349-450 with get_neg_points=False and
    num_points_for_sam == 1 (validation_protosam.py:226).  Returns (points
    [ncc,npts,2], labels [sum npts]) -- int64 for 'conf', float64 otherwise."""
    assert k == 1, "the oracle restates the production setting num_points_for_sam=1"
    ids = [int(i) for i in np.unique(cc[1]) if i != 0]
    _, pt, _ = _cc_prompts(p_fg, cc[1], int(cc[1].max()) + 1)
    pts = []
    for cid in ids:
        conf_pt = pt[cid][None, :]
        if point_mode == "conf":
            pts.append(conf_pt)
        elif point_mode == "centroid":
            pts.append(cc[3][cid][None, :])
        elif point_mode == "both":
            pts.append(np.vstack([conf_pt, cc[3][cid][None, :]]))
        else:
            raise NotImplementedError(f"point mode {point_mode} not implemented")
    labels = np.array([l + 1 for l, p in enumerate(pts) for _ in range(len(p))])
    return np.stack(pts), labels


def coarse_to_prompts(low_logits, mid_size, out_size=1024, use_cca=False, point_mode="both"):
    """ProtoSAM.forward lines 592-635 for one query slice and one label.

    Returns dict(pred uint8 [out,out], p_fg, n, labels, stats, centroids, conf,
    bboxes int64 [ncc,4], points, point_labels); ``empty`` True mirrors the
    early return at :612-613."""
    logits, p, pred = coarse_logits_to_probs(low_logits, mid_size, out_size)
    p_fg = p[0, 1]
    if use_cca:
        cc = cca(pred, p_fg, return_cc=True)
        conf = None
    else:
        cc, conf = get_connected_components(pred, p_fg, return_conf=True)
    out = dict(pred=pred, p_fg=p_fg, n=cc[0], labels=cc[1], stats=cc[2], centroids=cc[3], conf=conf,
               empty=bool(pred.max() == 0))
    if out["empty"]:
        return out
    out["bboxes"] = get_bbox_per_cc(cc)
    out["points"], out["point_labels"] = get_sam_input_points(cc, p_fg, point_mode)
    return out


# --------------------------------------------------------------------------
# F. ProtoMedSAM variant -- models/ProtoMedSAM.py:175-200
# --------------------------------------------------------------------------

def need_softmax(x, dim=1):
    """util/utils.py:62-63: ``not all(isclose(x.sum(dim), 1) & (x >= 0))`` (the [N,H,W] sum test broadcasts against
    the [N,2,H,W] sign test)."""
    x = np.asarray(x, dtype=np.float32)
    s = x.sum(axis=dim, dtype=np.float32)
    close = np.isclose(s, np.ones_like(s))          # torch.isclose defaults: rtol=1e-5, atol=1e-8, as numpy
    return not bool(np.all(np.expand_dims(close, dim) & (x >= 0)))


def coarse_to_prompts_medsam(low_logits, mid_size, out_size=1024, use_cca=False, image_size=(1024, 1024)):
    """ProtoMedSAM.forward lines 175-200 for one query slice and one label.

    The coarse logits are turned into probabilities FIRST when ``need_softmax`` says so (:178-179); the mask is
    their argmax; ``cca`` / ``get_connected_components`` then apply ``softmax(1)`` to those probabilities AGAIN
    (util/utils.py:486), so the per-component confidences -- and with ``use_cca`` the component kept -- are those of
    a softmax of a softmax.  Boxes are ``get_bbox_per_cc / [W,H,W,H] * max(image_size)`` in float64 (:197-198)."""
    logits, p, _ = coarse_logits_to_probs(low_logits, mid_size, out_size)
    output = p if need_softmax(logits) else logits
    pred = (output[0, 1] > output[0, 0]).astype(np.uint8)            # argmax over two classes keeps class 0 on ties
    conf_p = softmax2(output)[0, 1]
    if use_cca:
        cc = cca(pred, conf_p, return_cc=True)
        conf = None
    else:
        cc, conf = get_connected_components(pred, conf_p, return_conf=True)
    out = dict(pred=pred, conf_p=conf_p, n=cc[0], labels=cc[1], stats=cc[2], centroids=cc[3], conf=conf,
               empty=bool(pred.max() == 0), need_softmax=need_softmax(logits))
    if out["empty"]:
        return out
    H = W = out_size
    out["bboxes"] = get_bbox_per_cc(cc)
    out["boxes_1024"] = out["bboxes"] / np.array([W, H, W, H]) * max(image_size)
    return out


# --------------------------------------------------------------------------
# G. optional prompt variants -- models/ProtoSAM.py:361-434, 452-498, 580-590
# --------------------------------------------------------------------------

def get_confidence_from_logits(logits):
    """util/utils.py:429-434 (``coarse_pred_only``): mean foreground probability over the pixels with p >= 0.5.
    The reference sums in float32 (ATen's vectorised cascade); the sums here are float64, so compare at ~1e-6."""
    p = softmax2(np.asarray(logits, dtype=np.float32))[0, 1].astype(np.float64).ravel()
    sel = p >= 0.5
    return float(p[sel].sum() / (sel.sum() + 1e-6))


def dilate_square(mask, iterations=10):
    """cv2.dilate(mask, np.ones((3,3)), iterations=n): n passes of a 3x3 maximum = one (2n+1)^2 maximum clipped at the
    image border (Chebyshev distance <= n)."""
    m = np.asarray(mask).astype(bool)
    out = m.copy()
    for _ in range(iterations):
        nxt = out.copy()
        nxt[1:, :] |= out[:-1, :]; nxt[:-1, :] |= out[1:, :]
        out = nxt
        nxt = out.copy()
        nxt[:, 1:] |= out[:, :-1]; nxt[:, :-1] |= out[:, 1:]
        out = nxt
    return out


def get_most_conf_points(p_fg, pred, k):
    """ProtoSAM.get_most_conf_points (models/ProtoSAM.py:266-289) for any k: (locations int64 [k,2] in (x, y),
    [confidences]) = torch.nonzero(mask)[torch.topk(p_fg[mask], k).indices], torch.topk's order among equal values
    replayed (psamo_topk_pos); (None, None) for an empty mask."""
    L = lib()
    mask = np.asarray(pred).astype(bool)
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return None, None
    v = _f32(np.asarray(p_fg)[ys, xs])
    if k > len(v):
        raise RuntimeError("selected index k out of range")
    pos = np.zeros(k, np.int32)
    L.psamo_topk_pos.restype = ctypes.c_int
    rc = L.psamo_topk_pos(_ptr(v), int(len(v)), int(k), pos.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return np.stack([xs[pos], ys[pos]], 1).astype(np.int64), [float(c) for c in v[pos]]


def _topk1_masked(values, mask):
    """get_most_conf_points(values, mask, 1) (models/ProtoSAM.py:266-289): (x, y) of torch.topk(values[mask], 1), or
    None when the mask is empty."""
    L = lib()
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return None
    v = _f32(values[ys, xs])
    pos = L.psamo_topk1_pos(_ptr(v), int(len(v)))
    return np.array([[xs[pos], ys[pos]]], dtype=np.int64)


def get_neg_points(cc, output_p, l=1, host_aliasing=True, thresh=0.95, iterations=10):
    """The negative points of ``get_sam_input_points(..., get_neg_points=True)`` (models/ProtoSAM.py:361-434) for
    l = 1: per component, vstack([most confident background point in the ring of Chebyshev width 10 around the
    component, most confident background point of the image with p_bg >= 0.95]).

    ``host_aliasing``: on a CPU tensor ``output_p[0, 0].detach().cpu()`` is a VIEW, so the in-place thresholding at
    :364 (``bg_p[bg_p < 0.95] = 0``) also zeroes the map the ring search reads at :414 -- that is what the CPU-generated
    fixtures contain; on the reference's CUDA path ``.cpu()`` copies and the ring search sees the raw p_bg
    (host_aliasing=False)."""
    assert l == 1
    p_bg = _f32(output_p[0, 0])
    thr = np.where(p_bg < np.float32(thresh), np.float32(0), p_bg)
    glob = _topk1_masked(thr, thr > 0)
    ring_src = thr if host_aliasing else p_bg
    out = []
    for cid in [int(i) for i in np.unique(cc[1]) if i != 0]:
        comp = cc[1] == cid
        ring = dilate_square(comp, iterations) & ~comp
        neg = _topk1_masked(ring_src, ring)
        if neg is not None and glob is not None:
            neg = np.vstack([neg, glob])
        else:
            neg = glob if neg is None else neg
        out.append(neg)
    return out


def sam_mask_inputs(cc, size=256):
    """get_sam_input_mask + predict_w_masks (models/ProtoSAM.py:452-498): per component the 0/1 mask resized to
    256x256 with cv2.INTER_NEAREST (source index = min(floor(dst * src/dst), src-1)), foreground -> 10, background
    -> -8, handed to SamPredictor.predict as ``mask_input = in_mask[None].astype(np.uint8)`` (-8 wraps to 248).
    Returns (uint8 [ncc,1,size,size], labels [ncc])."""
    ids = [int(i) for i in np.unique(cc[1]) if i != 0]
    H, W = cc[1].shape
    # cv2 resizeNN: sx = min(cvFloor(x * ifx), src - 1) with ifx = 1 / inv_scale_x, inv_scale_x = (double)dst / src
    ys = np.minimum(np.floor(np.arange(size) * (1.0 / (size / H))).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(size) * (1.0 / (size / W))).astype(np.int64), W - 1)
    small = cc[1][ys[:, None], xs[None, :]]
    masks = np.stack([np.where(small == cid, 10, 248).astype(np.uint8)[None] for cid in ids]) if ids else \
        np.zeros((0, 1, size, size), np.uint8)
    return masks, np.array(ids)
