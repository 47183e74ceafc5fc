"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where the reference tree
is mounted at /root/reference:

    python -m oracle.make_golden            # rewrites tests/golden/

The reference holds no tests or golden vectors for this path,
so these fixtures -- outputs of the reference's own code on seeded inputs -- are
what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path
(tests/test_gpu_*.py, which also read the fixtures directly).

Reference entry points executed:
  * models/alpmodule.py  MultiProtoAsConv.forward / get_prototypes   (alp_*.npz)
  * models/ProtoSAM.py   ProtoSAM.forward lines 536-678 with a stub coarse model
    returning fixed logits and a capturing SamPredictor              (prompt_*.npz)
  * models/grid_proto_fewshot.py FewShotSeg.forward with a stub DINOv2 (fss_*.npz)
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from protosam_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _versions():
    import cv2
    return np.array([f"torch={torch.__version__}", f"numpy={np.__version__}", f"cv2={cv2.__version__}",
                     f"cpu_capability={torch.backends.cpu.get_cpu_capability()}"])


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------ ALP module

def run_ref_alp(qry, sup_x, sup_y, mode, thresh, proto_grid, feature_hw, isval, val_wsize, vis_sim=True):
    """Call the reference module; returns a dict of numpy outputs (or {'error': ...})."""
    alp = ref_shims.load_alpmodule()
    with _quiet():
        m = alp.MultiProtoAsConv(proto_grid=proto_grid, feature_hw=feature_hw)
    out = {}
    with ref_shims.cpu_cuda_identity(), _quiet(), torch.no_grad():
        S = sup_x.shape[1]
        sx = sup_x.squeeze(0).squeeze(1)
        sy = sup_y.squeeze(0).reshape(S, 1, sx.shape[-2], sx.shape[-1])
        vw = val_wsize if val_wsize is not None else m.avg_pool_op.kernel_size[0]
        pro_n, grid, nz = m.get_prototypes(sx, sy, mode, vw, thresh, isval)
        out["pro_n"] = pro_n.numpy().copy()
        out["non_zero"] = nz.numpy().copy()
        if mode != "mask":
            ks = (vw, vw) if isval else tuple(m.kernel_size)
            pooled = F.avg_pool2d(sy, ks)
            out["survive"] = (pooled.reshape(-1) > thresh).numpy()
        try:
            pred, assign, vis, pgrid = m(qry, sup_x, sup_y, mode, thresh, isval=isval, val_wsize=val_wsize,
                                         vis_sim=vis_sim)
        except RuntimeError as e:
            out["error"] = np.array(str(e)[:200])
            return out
        out["pred_grid"] = pred.numpy().copy()
        out["debug_assign"] = assign[0].numpy().copy()
        out["proto_assign"] = vis["proto_assign"].numpy().copy()
        if vis_sim:
            out["raw_local_sims"] = vis["raw_local_sims"].numpy().copy()
        out["proto_grid"] = pgrid.numpy().copy()
    return out


def alp_small_cases():
    """Small shapes, inputs stored.  Covers modes, multi-shot, odd sizes, window
    sizes, training-style (isval=False, kernel_size from proto_grid), soft masks,
    no-survivor and zero-prototype edge cases, 4-D and 5-D query."""
    cases = []
    rng_id = 0

    def mk(name, C, h, w, S, mode, isval, val_wsize, proto_grid, mask_kind="ellipse", thresh=0.95, q5d=True):
        nonlocal rng_id
        rng_id += 1
        seed = 9000 + rng_id
        sup = synth.layer_norm(synth.gaussian_like(seed, (S, h, w, C)))
        qry = synth.layer_norm(sup[0] + 0.5 * synth.gaussian_like(seed + 50, (h, w, C)))
        if mask_kind == "ellipse":
            y = np.stack([synth.nearest_resize(synth.ellipse_mask(seed + s, 8 * max(h, w), lo=0.15, hi=0.4), h, w)
                          for s in range(S)])
        elif mask_kind == "soft":
            y = synth.uniform(seed + 7, (S, h, w)) ** 0.15
        elif mask_kind == "empty":
            y = np.zeros((S, h, w), np.float32)
        elif mask_kind == "full":
            y = np.ones((S, h, w), np.float32)
        elif mask_kind == "tiny":
            y = np.zeros((S, h, w), np.float32); y[:, h // 2, w // 2] = 1
        else:
            raise ValueError(mask_kind)
        cases.append(dict(name=name, sup=sup, qry=qry, y=y.astype(np.float32), mode=mode, isval=isval,
                          val_wsize=val_wsize, proto_grid=proto_grid, thresh=thresh, q5d=q5d))

    for mode in ("mask", "gridconv", "gridconv+"):
        mk(f"s1_16_{mode}", 32, 16, 16, 1, mode, True, 2, [8, 8])
        mk(f"s3_16_{mode}", 32, 16, 16, 3, mode, True, 2, [8, 8])
        mk(f"s1_37_ws3_{mode}", 24, 37, 37, 1, mode, True, 3, [8, 8])
        mk(f"s2_21x29_train_{mode}", 16, 21, 29, 2, mode, False, None, [4, 4])       # kernel [5,7]
    for ws in (2, 4, 5, 8):
        mk(f"s1_32_ws{ws}_gridconv+", 40, 32, 32, 1, "gridconv+", True, ws, [8, 8])
        mk(f"s1_32_ws{ws}_gridconv", 40, 32, 32, 1, "gridconv", True, ws, [8, 8], q5d=False)
    mk("soft_gridconv+", 32, 16, 16, 2, "gridconv+", True, 2, [8, 8], mask_kind="soft")
    mk("soft_gridconv_t05", 32, 16, 16, 1, "gridconv", True, 4, [8, 8], mask_kind="soft", thresh=0.9)
    mk("empty_gridconv+", 32, 16, 16, 1, "gridconv+", True, 2, [8, 8], mask_kind="empty")
    mk("empty_gridconv_ERR", 32, 16, 16, 1, "gridconv", True, 2, [8, 8], mask_kind="empty")
    mk("empty_mask", 32, 16, 16, 1, "mask", True, 2, [8, 8], mask_kind="empty")
    mk("tiny_gridconv+", 32, 16, 16, 1, "gridconv+", True, 2, [8, 8], mask_kind="tiny")
    mk("full_gridconv", 32, 16, 16, 1, "gridconv", True, 2, [8, 8], mask_kind="full")
    mk("train_default_ks", 32, 32, 32, 1, "gridconv+", False, None, [8, 8])          # kernel [4,4]
    return cases


def gen_alp_small():
    store = {"versions": _versions()}
    names = []
    for c in alp_small_cases():
        S, h, w, C = c["sup"].shape
        # channels-last physical layout, logical [1,S,1,C,h,w] / [1,1,C,h,w] like the caller builds them
        sup_x = torch.from_numpy(c["sup"]).permute(0, 3, 1, 2)[None, :, None]
        qry = torch.from_numpy(c["qry"]).permute(2, 0, 1)[None, None]
        if not c["q5d"]:
            qry = qry[:, 0]
        sup_y = torch.from_numpy(c["y"])[None, :, None]
        out = run_ref_alp(qry, sup_x, sup_y, c["mode"], c["thresh"], c["proto_grid"], [h, w], c["isval"],
                          c["val_wsize"])
        n = c["name"]
        names.append(n)
        store[f"{n}/sup"] = c["sup"]; store[f"{n}/qry"] = c["qry"]; store[f"{n}/y"] = c["y"]
        store[f"{n}/meta"] = np.array([c["mode"], str(int(c["isval"])), str(c["val_wsize"]),
                                       str(c["proto_grid"][0]), str(c["thresh"]), str(int(c["q5d"]))])
        for k, v in out.items():
            store[f"{n}/{k}"] = v
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "alp_small.npz"), **store)
    print("alp_small:", len(names), "cases")


def gen_alp_config_shapes():
    """Real config shapes; inputs are regenerated from synth by the tests (not stored)."""
    store = {"versions": _versions()}
    names = []
    for cfg_name, nq in (("cfg1_vits_256", 1), ("cfg2_chaos_mri", 2)):
        cfg = synth.CONFIGS[cfg_name]
        L = min(cfg["L"], 2)
        vol = synth.make_volume(1234, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
        sup_x = torch.from_numpy(vol.sup).permute(0, 3, 1, 2)[None, :, None]
        for l in range(L):
            for q in range(nq):
                qry = torch.from_numpy(vol.qry[q]).permute(2, 0, 1)[None, None]
                for kind, mask, mode in (("bg", vol.bg[l], "gridconv"), ("fg", vol.fg[l], "gridconv+"),
                                         ("fgmask", vol.fg[l], "mask")):
                    sup_y = torch.from_numpy(mask)[None, :, None]
                    out = run_ref_alp(qry, sup_x, sup_y, mode, 0.95, [8, 8], [cfg["h"], cfg["w"]], True, cfg["ws"],
                                      vis_sim=False)
                    n = f"{cfg_name}/l{l}/q{q}/{kind}"
                    names.append(n)
                    for k in ("pred_grid", "debug_assign", "survive", "pro_n", "proto_grid"):
                        if k in out:
                            store[f"{n}/{k}"] = out[k]
        store[f"{cfg_name}/meta"] = np.array([1234, nq, L])
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "alp_configs.npz"), **store)
    print("alp_configs:", len(names), "cases")


def gen_alp_config_shapes2():
    """The remaining BASELINE shapes (configs 3, 4 and the window sweep of config 5 at C = 1024), one query slice each.
    Only the small outputs are stored (maps, survival masks, assignments): the prototype rows of these shapes would
    add tens of MB to the repository and are already pinned at C <= 768 by alp_configs.npz."""
    store = {"versions": _versions()}
    names = []
    jobs = []      # (key, cfg-like dict, seed, L, ws, [(kind, mode)])
    c3 = synth.CONFIGS["cfg3_synapse_ct"]
    jobs.append(("cfg3_synapse_ct", c3, 1234, 2, c3["ws"], [("bg", "gridconv"), ("fg", "gridconv+"), ("fgmask", "mask")]))
    c4 = synth.CONFIGS["cfg4_polyp_1024"]
    jobs.append(("cfg4_polyp_1024", c4, 1234, 1, c4["ws"], [("bg", "gridconv"), ("fg", "gridconv+"), ("fgmask", "mask")]))
    c5 = synth.CONFIGS["cfg5_stress_vitl"]
    for ws in (3, 6, 7):
        jobs.append((f"cfg5_ws{ws}", c5, 50 + ws, 1, ws, [("bg", "gridconv"), ("fg", "gridconv+")]))
    for key, cfg, seed, L, ws, kinds in jobs:
        vol = synth.make_volume(seed, Q=1, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
        sup_x = torch.from_numpy(vol.sup).permute(0, 3, 1, 2)[None, :, None]
        qry = torch.from_numpy(vol.qry[0]).permute(2, 0, 1)[None, None]
        for l in range(L):
            for kind, mode in kinds:
                mask = vol.bg[l] if kind == "bg" else vol.fg[l]
                out = run_ref_alp(qry, sup_x, torch.from_numpy(mask)[None, :, None], mode, 0.95, [8, 8],
                                  [cfg["h"], cfg["w"]], True, ws, vis_sim=False)
                n = f"{key}/l{l}/q0/{kind}"
                names.append(n)
                for k in ("pred_grid", "debug_assign", "survive"):
                    if k in out:
                        store[f"{n}/{k}"] = out[k]
        store[f"{key}/meta"] = np.array([seed, 1, L, ws])
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "alp_configs2.npz"), **store)
    print("alp_configs2:", len(names), "cases")


# ------------------------------------------------------------------ prompts

def run_ref_protosam(logits_S, use_cca, point_mode, S):
    """ProtoSAM.forward with a stub coarse model returning ``logits_S`` [1,2,S,S]
    (what FewShotSeg returns) and the capturing predictor."""
    _, PS, _ = ref_shims.load_pipeline()
    with _quiet():
        model = PS.ProtoSAM(image_size=(1024, 1024),
                            coarse_segmentation_model=ref_shims.FixedLogitsCoarseModel(logits_S),
                            num_points_for_sam=1, use_points=True, use_bbox=True, use_cca=use_cca,
                            point_mode=point_mode)
    model.eval()
    img = torch.from_numpy(synth.uniform(77, (1, 3, S, S)))
    with torch.no_grad(), _quiet():
        pred, scores = model(img, ref_shims._NullInput(), degrees_rotate=0)
    return model.predictor.calls, pred


def smooth_field(seed, h, w, amp=18.0, cells=5):
    """Random smooth 2-channel logits (bilinear-upsampled noise) in about +-amp."""
    lo = torch.from_numpy(synth.gaussian_like(seed, (1, 2, cells, cells)))
    x = F.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False)
    return (x * amp / 2).numpy().astype(np.float32)


def prompt_cases():
    cases = []
    # (name, low logits [1,2,h,w], S)
    for i, (h, S) in enumerate(((32, 256), (37, 518), (48, 672), (73, 1024))):
        cases.append((f"smooth_{h}_{S}", smooth_field(100 + i, h, h), S))
        sp = synth.gaussian_like(200 + i, (1, 2, h, h)) * 6.0            # speckle: many components
        cases.append((f"speckle_{h}_{S}", sp.astype(np.float32), S))
    sat = smooth_field(300, 37, 37, amp=60.0)                             # saturating probabilities (ties at 1.0)
    cases.append(("saturated_37_518", sat, 518))
    empty = np.stack([np.full((1, 24, 24), 3.0, np.float32), np.full((1, 24, 24), -3.0, np.float32)], 1)
    cases.append(("empty_24_256", empty.astype(np.float32), 256))
    full = -empty
    cases.append(("full_24_256", full.astype(np.float32), 256))
    blobs = np.full((1, 2, 32, 32), 0.0, np.float32)
    blobs[0, 0] = 2.0
    for (cy, cx, r, a) in ((6, 6, 3, 9.0), (20, 24, 5, 12.0), (26, 6, 2, 30.0), (10, 22, 1, 5.0)):
        yy, xx = np.mgrid[0:32, 0:32]
        blobs[0, 1][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = a
    cases.append(("blobs_32_256", blobs, 256))
    return cases


def gen_prompts():
    store = {"versions": _versions()}
    names = []
    for name, low, S in prompt_cases():
        low_t = torch.from_numpy(low)
        logits_S = F.interpolate(low_t, size=(S, S), mode="bilinear")        # grid_proto_fewshot.py:272-273
        store[f"{name}/low"] = low
        store[f"{name}/S"] = np.array(S)
        # the 1024^2 mask the reference derives (ProtoSAM.py:592-602), bit-packed, for debugging parity
        lg = F.interpolate(logits_S, size=(1024, 1024), mode="bilinear") if S != 1024 else logits_S
        p = lg.softmax(1)
        pred = p.argmax(1)[0].numpy().astype(np.uint8)
        store[f"{name}/pred_bits"] = np.packbits(pred)
        store[f"{name}/p_fg_sample"] = p[0, 1, ::61, ::67].numpy().copy()
        for use_cca in (False, True):
            for pm in ("conf", "centroid", "both"):
                if name.startswith("speckle") and not (pm == "both"):
                    continue                                                    # keep fixture + runtime small
                calls, out_pred = run_ref_protosam(logits_S, use_cca, pm, S)
                key = f"{name}/cca{int(use_cca)}_{pm}"
                names.append(key)
                store[f"{key}/ncalls"] = np.array(len(calls))
                if calls:
                    store[f"{key}/points"] = np.stack([c["point_coords"] for c in calls])
                    store[f"{key}/point_labels"] = np.stack([c["point_labels"] for c in calls])
                    store[f"{key}/boxes"] = np.stack([c["box"] for c in calls])
                    store[f"{key}/multimask"] = np.array([bool(c["multimask_output"]) for c in calls])
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "prompts.npz"), **store)
    print("prompts:", len(names), "cases")


# ------------------------------------------------------------------ optional variants

def variant_cases():
    cases = {n: (low, S) for n, low, S in prompt_cases()}
    return [(n,) + cases[n] for n in ("blobs_32_256", "smooth_32_256", "smooth_37_518", "speckle_32_256",
                                      "saturated_37_518", "empty_24_256")]


def gen_variants():
    """Negative points, mask prompts, coarse_pred_only confidence (models/ProtoSAM.py:361-434, 452-498, 580-590) and
    the ProtoMedSAM box path (models/ProtoMedSAM.py:175-200), all through the reference's own forward()."""
    _, PS, uu = ref_shims.load_pipeline()
    PM = ref_shims.load_protomedsam()
    store = {"versions": _versions()}
    names = []
    for name, low, S in variant_cases():
        low_t = torch.from_numpy(low)
        logits_S = F.interpolate(low_t, size=(S, S), mode="bilinear")
        store[f"{name}/low"] = low
        store[f"{name}/S"] = np.array(S)
        names.append(name)
        img = torch.from_numpy(synth.uniform(77, (1, 3, S, S)))

        def run(device_semantics=False, **kw):
            with _quiet():
                model = PS.ProtoSAM(image_size=(1024, 1024), coarse_segmentation_model=ref_shims.FixedLogitsCoarseModel(logits_S),
                                    num_points_for_sam=1, **kw)
            model.eval()
            ctx = ref_shims.device_copy_semantics() if device_semantics else contextlib.nullcontext()
            with torch.no_grad(), _quiet(), ctx:
                pred, scores = model(img, ref_shims._NullInput(), degrees_rotate=0)
            return model.predictor.calls, pred, scores

        # negative points, both .cpu() semantics
        for dev_sem in (False, True):
            for use_cca in (False, True):
                calls, _, _ = run(dev_sem, use_points=True, use_bbox=True, use_cca=use_cca, point_mode="both",
                                  use_neg_points=True)
                key = f"{name}/neg_dev{int(dev_sem)}_cca{int(use_cca)}"
                store[f"{key}/ncalls"] = np.array(len(calls))
                if calls:
                    # every call has 2 positive + up to 2 negative points; pad to 4 rows with -1
                    pts = np.full((len(calls), 4, 2), -1.0)
                    lab = np.full((len(calls), 4), -1, np.int64)
                    for i, c in enumerate(calls):
                        pts[i, : len(c["point_coords"])] = c["point_coords"]
                        lab[i, : len(c["point_labels"])] = c["point_labels"]
                    store[f"{key}/points"] = pts
                    store[f"{key}/point_labels"] = lab
        # mask prompts
        calls, _, _ = run(False, use_points=False, use_bbox=False, use_mask=True, use_cca=False, point_mode="both")
        mcalls = [c for c in calls if c.get("mask_input") is not None]
        store[f"{name}/mask/ncalls"] = np.array(len(mcalls))
        if mcalls:
            m = np.stack([c["mask_input"] for c in mcalls])              # [ncc,1,256,256] uint8
            store[f"{name}/mask/values"] = np.unique(m)
            store[f"{name}/mask/fg_bits"] = np.packbits(m == 10)
            store[f"{name}/mask/dtype"] = np.array(str(m.dtype))
        # coarse_pred_only: confidence (+ cca)
        for use_cca in (False, True):
            _, pred, scores = run(False, use_points=True, use_bbox=True, use_cca=use_cca, point_mode="both",
                                  coarse_pred_only=True)
            store[f"{name}/coarse_cca{int(use_cca)}/conf"] = np.array(float(scores[0]))
            store[f"{name}/coarse_cca{int(use_cca)}/pred_bits"] = np.packbits(np.asarray(pred).astype(np.uint8))
        # ProtoMedSAM: boxes handed to medsam_inference, and the confidences its cca() sees
        for use_cca in (False, True):
            with _quiet():
                med = PM.ProtoMedSAM(image_size=(1024, 1024),
                                     coarse_segmentation_model=ref_shims.FixedLogitsCoarseModel(logits_S), use_cca=use_cca)
            med.eval()
            with torch.no_grad(), _quiet():
                med(img, ref_shims._NullInput(), degrees_rotate=0)
            key = f"{name}/medsam_cca{int(use_cca)}"
            store[f"{key}/ncalls"] = np.array(len(med.captured_boxes))
            if med.captured_boxes:
                store[f"{key}/boxes"] = med.captured_boxes[0]
        lg = F.interpolate(logits_S, size=(1024, 1024), mode="bilinear")
        P = lg.softmax(1)                                                # what ProtoMedSAM hands to cca() after need_softmax
        store[f"{name}/medsam_need_softmax"] = np.array(bool(uu.need_softmax(lg)))
        pred = np.array(P.argmax(1)[0])
        _, conf = uu.get_connected_components(pred, P, return_conf=True)
        store[f"{name}/medsam_conf"] = np.array([float(conf[k]) for k in sorted(conf)])
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "variants.npz"), **store)
    print("variants:", len(names), "cases")


def gen_topk():
    """ProtoSAM.get_most_conf_points(output_p_fg, pred, k) for k > 1 (models/ProtoSAM.py:266-289), called as the
    reference defines it, on the 1024^2 probability maps of the prompt cases: per case and k, the first components (cv2
    label order) that have at least k pixels."""
    _, PS, uu = ref_shims.load_pipeline()
    store = {"versions": _versions()}
    names = []
    for name, low, S in variant_cases():
        logits_S = F.interpolate(torch.from_numpy(low), size=(S, S), mode="bilinear")
        lg = F.interpolate(logits_S, size=(1024, 1024), mode="bilinear")
        P = lg.softmax(1)
        pred = np.array(P.argmax(1)[0])
        cc, _ = uu.get_connected_components(pred, lg)
        if cc[0] <= 1:
            continue
        names.append(name)
        store[f"{name}/low"] = low
        store[f"{name}/S"] = np.array(S)
        for k in (2, 5, 17):
            labs, locs, confs = [], [], []
            for j in range(1, cc[0]):
                if cc[2][j, 4] < k or len(labs) >= 6:
                    continue
                loc, conf = PS.ProtoSAM.get_most_conf_points(None, P[0, 1], torch.from_numpy(cc[1] == j), k)
                labs.append(j); locs.append(loc); confs.append(np.array(conf, np.float64))
            store[f"{name}/k{k}/labels"] = np.array(labs, np.int64)
            if labs:
                store[f"{name}/k{k}/locations"] = np.stack(locs)
                store[f"{name}/k{k}/confidences"] = np.stack(confs)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "topk.npz"), **store)
    print("topk:", len(names), "cases")


def main():
    assert ref_shims.reference_available(), "reference tree not mounted"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(1234)
    gen_alp_small()
    gen_alp_config_shapes()
    gen_alp_config_shapes2()
    gen_prompts()
    gen_variants()
    gen_topk()


if __name__ == "__main__":
    main()
