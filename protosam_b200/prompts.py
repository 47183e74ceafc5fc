"""Host-side mirror of the reference's coarse-mask -> SAM-prompt helpers.

The reference runs this stage on the CPU with OpenCV/numpy (util/utils.py:474-541,
models/ProtoSAM.py:242-289, 349-450, 592-635).  Here kernel 3 (libpsam_b200.so) does the work on
the GPU and emits compact records; this module only re-packages those records into the exact
objects the reference hands to ``SamPredictor.predict`` (models/ProtoSAM.py:500-533): same names,
argument meaning, dtypes and shapes.  No arithmetic on the maps happens in Python.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _lib, ops

CONF_MODE, CENTROID_MODE, BOTH_MODE = "conf", "centroid", "both"
POINT_MODES = (CONF_MODE, CENTROID_MODE, BOTH_MODE)


class ConnectedComponents(tuple):
    """The cv2-style 4-tuple ``(n, labels, stats, centroids)`` the reference's helpers pass around
    (util/utils.py:478), carrying the GPU records it was built from."""

    def __new__(cls, n, labels, stats, centroids, records=None, header=None, dev=None):
        self = super().__new__(cls, (n, labels, stats, centroids))
        self.records = records
        self.header = header
        self.dev = dev          # device tensors the optional variants (negative points, mask prompts) work on
        return self


@dataclass
class SlicePrompts:
    """Everything ProtoSAM.forward derives for one (query slice, label) before calling SAM."""
    empty: bool                      # ProtoSAM.py:612-613 early return
    ncc: int                         # components found
    boxes: Optional[np.ndarray]      # int64 [n,4] XYXY           get_bbox_per_cc
    points: Optional[np.ndarray]     # [n,npts,2] int64|float64   get_sam_input_points
    point_labels: Optional[np.ndarray]
    conf: Optional[dict]             # {label: confidence}, None with use_cca
    multimask_output: bool           # ProtoSAM.py:522
    flags: int
    neg_points: Optional[list] = None    # per component int64 [k,2] (k <= 2) or None   use_neg_points
    mask_inputs: Optional[np.ndarray] = None   # uint8 [n,1,256,256]                       use_mask

    def predict_calls(self) -> List[dict]:
        """kwargs of each SamPredictor.predict call, as predict_w_points_bbox builds them (:505-523)."""
        if self.empty:
            return []
        calls = []
        for i in range(len(self.boxes)):
            pts, lab = self.points[i], np.array([1] * len(self.points[i]))
            if self.neg_points is not None:                      # :509-512
                neg = self.neg_points[i]
                neg = [] if neg is None else [q for q in neg]
                pts = np.vstack([pts, *neg]) if neg else np.vstack([pts])
                lab = np.array([1] * len(self.points[i]) + [0] * len(neg))
            calls.append(dict(point_coords=pts, point_labels=lab, box=self.boxes[i], multimask_output=self.multimask_output))
        return calls

    def mask_predict_calls(self) -> List[dict]:
        """kwargs of each SamPredictor.predict call of predict_w_masks (:477-480)."""
        if self.empty or self.mask_inputs is None:
            return []
        return [dict(mask_input=m, multimask_output=True) for m in self.mask_inputs]


def _points_from_records(recs: np.ndarray, point_mode: str) -> np.ndarray:
    if point_mode == CONF_MODE:
        return recs["conf_pt"][:, None, :].astype(np.int64)                       # [n,1,2] int64
    if point_mode == CENTROID_MODE:
        return recs["centroid"][:, None, :].astype(np.float64)                    # [n,1,2] float64
    if point_mode == BOTH_MODE:                                                   # np.vstack -> float64
        return np.stack([recs["conf_pt"].astype(np.float64), recs["centroid"]], axis=1)
    raise NotImplementedError(f"point mode {point_mode} not implemented")


def prompts_from_records(hdr: np.void, recs: np.ndarray, use_cca: bool, point_mode: str = BOTH_MODE) -> SlicePrompts:
    """One image's header + records -> the reference's prompt arrays."""
    if point_mode not in POINT_MODES:
        raise ValueError(f"point mode must be one of {POINT_MODES}")
    flags = int(hdr["flags"])
    if flags & _lib.IMG_RUN_OVERFLOW:
        raise RuntimeError("coarse mask has more foreground runs than the workspace holds; "
                           "re-run with a larger max_runs")
    if flags & _lib.IMG_CC_TRUNCATED:
        # the reference prompts SAM for EVERY component: a truncated list would silently drop prompts
        raise RuntimeError(f"coarse mask has {int(hdr['ncc'])} components, more than the max_cc={len(recs)} records "
                           "per image; re-run with a larger max_cc")
    if flags & _lib.IMG_CCA_AMBIGUOUS:
        import warnings
        warnings.warn("use_cca: two components' confidences differ by less than float32 summation error; the exact-sum "
                      "winner was kept (the reference's float32 numpy sum may pick the other)", RuntimeWarning)
    if flags & _lib.IMG_EMPTY:
        return SlicePrompts(True, 0, None, None, None, None, not use_cca, flags)
    n = int(hdr["n_rec"])
    r = recs[:n]
    pts = _points_from_records(r, point_mode)
    labels = np.array([l + 1 for l, p in enumerate(pts) for _ in range(len(p))])
    conf = None
    if not use_cca:
        conf = {0: 0}
        conf.update({int(lab): float(c) for lab, c in zip(r["label"], r["conf"])})
    return SlicePrompts(False, int(hdr["ncc"]), r["box"].astype(np.int64).copy(), pts, labels, conf,
                        not use_cca, flags)


def medsam_boxes(boxes_xyxy: np.ndarray, W: int, H: int, image_size=(1024, 1024)) -> np.ndarray:
    """ProtoMedSAM.forward's box hand-off (models/ProtoMedSAM.py:197-200): the per-component XYXY boxes of
    get_bbox_per_cc rescaled to MedSAM's input frame, ``bbox / [W, H, W, H] * max(image_size)`` in float64."""
    return boxes_xyxy / np.array([W, H, W, H]) * max(image_size)


def _neg_lists(neg_u8: torch.Tensor, n_rec: np.ndarray) -> List[list]:
    """device psam_neg_point records -> per image, per component: int64 [k,2] = vstack([ring point, global point]) with
    the missing ones dropped, or None (models/ProtoSAM.py:419-423)."""
    a = neg_u8.detach().cpu().numpy()
    N = np.frombuffer(a.tobytes(), dtype=ops.NEG_DTYPE).reshape(a.shape[0], a.shape[1])
    out = []
    for i in range(len(N)):
        g = N[i, -1]
        per = []
        for r in range(int(n_rec[i])):
            rows = [N[i, r]["pt"]] if N[i, r]["has"] else []
            if g["has"]:
                rows.append(g["pt"])
            per.append(np.stack(rows).astype(np.int64) if rows else None)
        out.append(per)
    return out


def coarse_to_prompts(low_logits: torch.Tensor, mid_size: int, out_size: int = 1024, use_cca: bool = False,
                      point_mode: str = BOTH_MODE, max_cc: int = ops.DEFAULT_MAX_CC,
                      max_runs: int = ops.DEFAULT_MAX_RUNS, use_neg_points: bool = False, use_mask: bool = False,
                      host_aliasing: bool = False, variant: str = "protosam") -> List[SlicePrompts]:
    """[n,2,h,w] coarse scores (CUDA) -> prompts per image.  Mirrors FewShotSeg's upsample
    (grid_proto_fewshot.py:270-273) + ProtoSAM.forward lines 592-635.

    use_neg_points / use_mask (models/ProtoSAM.py:361-434, 452-498; off in the reference's configs) take the every-pixel
    variant of kernel 3a: they need the background probabilities and the label image.  host_aliasing selects which of
    the reference's two behaviours the ring search reproduces (see psam_neg_points in include/psam_b200.h)."""
    prob_mode = "softmax_twice" if variant == "medsam" else "softmax"
    if not (use_neg_points or use_mask):
        hdr, recs = ops.coarse_to_prompts(low_logits, mid_size, out_size, use_cca, max_cc, max_runs, prob_mode=prob_mode)
        H, R = ops.decode_headers(hdr), ops.decode_records(recs)       # one D2H of ~n*(64+96*max_cc) bytes
        return [prompts_from_records(H[i], R[i], use_cca, point_mode) for i in range(len(H))]
    p_fg, bits, probs2 = ops.upsample_softmax(low_logits, mid_size, out_size, want_probs2=True, prob_mode=prob_mode)
    hdr, recs, labels = ops.components(bits, p_fg, use_cca, max_cc, max_runs, want_labels=True)
    H, R = ops.decode_headers(hdr), ops.decode_records(recs)
    out = [prompts_from_records(H[i], R[i], use_cca, point_mode) for i in range(len(H))]
    if use_neg_points:
        neg = _neg_lists(ops.neg_points(labels, probs2[:, 0], hdr, recs, use_cca=use_cca, host_aliasing=host_aliasing),
                         H["n_rec"])
        for sp, ng in zip(out, neg):
            sp.neg_points = None if sp.empty else ng
    if use_mask:
        masks, offsets = ops.mask_prompts(labels, hdr, recs, use_cca=use_cca, capacity=max(int(H["n_rec"].sum()), 1))
        M, off = masks.cpu().numpy(), offsets.cpu().numpy()
        for i, sp in enumerate(out):
            sp.mask_inputs = None if sp.empty else M[off[i]: off[i + 1], None]
    return out


# ---------------------------------------------------------------------------------------------
# Function-level drop-ins with the reference's names and return conventions.
# ---------------------------------------------------------------------------------------------

def _cc_from_full_logits(query_pred_logits: torch.Tensor, use_cca: bool, max_cc: int, max_runs: int):
    ops._need_cuda(query_pred_logits)
    assert query_pred_logits.dim() == 4 and query_pred_logits.shape[0] == 1 and query_pred_logits.shape[1] == 2
    S = query_pred_logits.shape[-1]
    p_fg, bits, probs2 = ops.upsample_softmax(query_pred_logits, S, S, want_probs2=True)
    hdr, recs, labels = ops.components(bits, p_fg, use_cca, max_cc, max_runs, want_labels=True)
    H, R = ops.decode_headers(hdr)[0], ops.decode_records(recs)[0]
    if int(H["flags"]) & _lib.IMG_RUN_OVERFLOW:
        raise RuntimeError("more foreground runs than max_runs")
    if int(H["flags"]) & _lib.IMG_CC_TRUNCATED:
        raise RuntimeError(f"{int(H['ncc'])} components exceed max_cc={max_cc}")
    r = R[: int(H["n_rec"])]
    n = len(r) + 1
    stats = np.zeros((n, 5), np.int32)
    cent = np.zeros((n, 2), np.float64)
    stats[0] = H["bg_stats"]
    cent[0] = H["bg_centroid"]
    stats[1:, 0] = r["box"][:, 0]
    stats[1:, 1] = r["box"][:, 1]
    stats[1:, 2] = r["box"][:, 2] - r["box"][:, 0] + 1
    stats[1:, 3] = r["box"][:, 3] - r["box"][:, 1] + 1
    stats[1:, 4] = r["area"]
    cent[1:] = r["centroid"]
    return ConnectedComponents(n, labels[0].cpu().numpy(), stats, cent, records=r, header=H,
                               dev=dict(labels=labels, probs2=probs2, hdr=hdr, recs=recs, use_cca=bool(use_cca)))


def get_connected_components(query_pred_original, query_pred_logits, return_conf=False,
                             max_cc=4096, max_runs=ops.DEFAULT_MAX_RUNS):
    """util/utils.py:474-494.  ``query_pred_logits`` [1,2,H,W] CUDA logits at full resolution; the
    mask is re-derived from them on the device (identical to ``query_pred_original`` by
    construction: it is their argmax)."""
    cc = _cc_from_full_logits(query_pred_logits, False, max_cc, max_runs)
    if not return_conf:
        return cc, None
    conf = {0: 0}
    conf.update({int(l): float(c) for l, c in zip(cc.records["label"], cc.records["conf"])})
    return cc, conf


def cca(query_pred_original, query_pred_logits, return_conf=False, return_cc=False,
        max_runs=ops.DEFAULT_MAX_RUNS):
    """util/utils.py:496-541: keep the most confident component."""
    cc = _cc_from_full_logits(query_pred_logits, True, 1, max_runs)
    if return_cc:
        return cc
    pred = np.asarray(query_pred_original.detach().cpu() if torch.is_tensor(query_pred_original)
                      else query_pred_original)
    if cc[0] < 2:
        out, max_conf = np.zeros_like(pred), 0
    else:
        out, max_conf = pred * (cc[1] == 1).astype(np.uint8), float(cc.records["conf"][0])
    return (out, max_conf) if return_conf else out


def get_bbox_per_cc(conn_components: ConnectedComponents) -> np.ndarray:
    """models/ProtoSAM.py:242-264 -> int64 [n-1,4] XYXY inclusive."""
    return conn_components.records["box"].astype(np.int64).copy()


def get_most_conf_points(conn_components: ConnectedComponents, cc_id: int, k: int = 1):
    """models/ProtoSAM.py:266-289 for one component: ``(locations int64 [k,2] in (x, y), [confidences])``.
    The reference masks the 1024^2 probability map on the host and calls torch.topk; here kernel 3b already found,
    per component, the highest p_fg with torch.topk's tie rule, so for k = 1 (production: num_points_for_sam = 1,
    validation_protosam.py:226) the answer is read off the component's record.  k > 1 replays torch.topk on the device
    over the component's pixels (psam_topk_points), equal probabilities in the reference's order; like torch.topk it
    raises when the component has fewer than k pixels."""
    recs = conn_components.records
    hit = np.nonzero(recs["label"] == cc_id)[0]
    if len(hit) == 0:
        return None, None
    r = recs[hit[0]]
    if k == 1:
        return r["conf_pt"].astype(np.int64)[None, :].copy(), [float(r["conf_pt_p"])]
    if int(r["area"]) < k:
        raise RuntimeError(f"selected index k out of range: component {cc_id} has {int(r['area'])} pixels, k = {k}")
    d = conn_components.dev
    cache = d.setdefault("topk", {})
    if k not in cache:                                   # one launch serves every component of the image
        pts, conf = ops.topk_points(d["labels"], d["probs2"][:, 1], d["hdr"], d["recs"], k, use_cca=d["use_cca"])
        cache[k] = (pts[0].cpu().numpy(), conf[0].cpu().numpy())
    pts, conf = cache[k]
    return pts[hit[0]].copy(), [float(c) for c in conf[hit[0]]]


def get_sam_input_points(conn_components: ConnectedComponents, output_p=None, get_neg_points=False, l=1,
                         point_mode=BOTH_MODE, host_aliasing=False):
    """models/ProtoSAM.py:349-450 (num_points_for_sam = 1).  With get_neg_points every component also gets
    vstack([most confident background point of its 10-pixel ring, most confident background point of the image]) (l = 1),
    found on the device from the probabilities the components were built from (``output_p`` is not re-read)."""
    pts = _points_from_records(conn_components.records, point_mode)
    labels = np.array([k + 1 for k, p in enumerate(pts) for _ in range(len(p))])
    if not get_neg_points:
        neg = [None for _ in range(len(pts))]
        return pts, labels, neg, np.array([0] * len(neg))
    if l != 1:
        raise NotImplementedError("only l = 1 (the reference's call, models/ProtoSAM.py:624) is computed on the device")
    d = conn_components.dev
    neg_u8 = ops.neg_points(d["labels"], d["probs2"][:, 0], d["hdr"], d["recs"], use_cca=d["use_cca"],
                            host_aliasing=host_aliasing)
    neg = _neg_lists(neg_u8, np.array([len(pts)]))[0]
    return pts, labels, neg, np.array([0] * len(neg))


def get_sam_input_mask(conn_components: ConnectedComponents):
    """models/ProtoSAM.py:452-466: (float masks [ncc,H,W], component ids)."""
    ids = [int(i) for i in np.unique(conn_components[1]) if i != 0]
    return np.stack([(conn_components[1] == i).astype(np.float32) for i in ids]), np.array(ids)


def sam_mask_inputs(conn_components: ConnectedComponents, size: int = 256) -> np.ndarray:
    """What predict_w_masks hands to SamPredictor.predict for every component (models/ProtoSAM.py:471-480):
    uint8 [ncc,1,size,size], nearest-resized on the device, 10 inside / 248 (= uint8(-8)) outside."""
    d = conn_components.dev
    n = len(conn_components.records)
    masks, _ = ops.mask_prompts(d["labels"], d["hdr"], d["recs"], use_cca=d["use_cca"], size=size, capacity=max(n, 1))
    return masks[:n, None].cpu().numpy()


def get_confidence_from_logits(query_pred_logits: torch.Tensor) -> float:
    """util/utils.py:429-434 for [1,2,H,W] CUDA logits (ProtoSAM's coarse_pred_only path, models/ProtoSAM.py:580-590)."""
    ops._need_cuda(query_pred_logits)
    assert query_pred_logits.dim() == 4 and query_pred_logits.shape[0] == 1 and query_pred_logits.shape[1] == 2
    S = query_pred_logits.shape[-1]
    p_fg, _, _ = ops.upsample_softmax(query_pred_logits, S, S)
    return float(ops.confidence(p_fg)[0].item())
