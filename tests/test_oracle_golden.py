"""The CPU oracle against fixtures produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from protosam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAP_TOL = 1e-4          # fp32 maps: north_star allows 1e-3; the oracle itself sits well inside


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _alp_small_names():
    return list(_load("alp_small.npz")["names"])


@pytest.mark.parametrize("n", _alp_small_names())
def test_alp_small(n):
    g = _load("alp_small.npz")
    mode, isval, vw, pg, thresh, q5d = g[f"{n}/meta"]
    isval, vw, pg, thresh = bool(int(isval)), (None if vw == "None" else int(vw)), int(pg), float(thresh)
    sup, qry, y = g[f"{n}/sup"], g[f"{n}/qry"], g[f"{n}/y"]
    S, h, w, C = sup.shape
    sup_x = np.transpose(sup, (0, 3, 1, 2))[None, :, None]          # channels-last storage, logical NCHW
    q = np.transpose(qry, (2, 0, 1))[None, None]
    if not int(q5d):
        q = q[:, 0]
    ks = [h // pg, w // pg]
    if f"{n}/error" in g.files:                                      # zero prototypes in 'gridconv'
        with pytest.raises(RuntimeError):
            O.alp_forward(q, sup_x, y[None, :, None], mode, thresh, ks, isval=isval, val_wsize=vw)
        return
    pred, assign, vis, grid = O.alp_forward(q, sup_x, y[None, :, None], mode, thresh, ks, isval=isval,
                                            val_wsize=vw, vis_sim=True)
    assert pred.shape == g[f"{n}/pred_grid"].shape
    np.testing.assert_allclose(pred, g[f"{n}/pred_grid"], atol=MAP_TOL, rtol=0)
    np.testing.assert_allclose(vis["raw_local_sims"], g[f"{n}/raw_local_sims"], atol=MAP_TOL, rtol=0)
    assert grid.shape == g[f"{n}/proto_grid"].shape and np.array_equal(grid, g[f"{n}/proto_grid"])
    if mode == "mask":
        np.testing.assert_allclose(assign[0], g[f"{n}/debug_assign"], atol=MAP_TOL, rtol=0)
        return
    assert np.array_equal(assign[0], g[f"{n}/debug_assign"])
    kk = (vw, vw) if isval else ks
    pr = O.get_prototypes(np.transpose(sup, (0, 3, 1, 2)), y[:, None], mode, kk, thresh, vw)
    assert np.array_equal(pr["survive"], g[f"{n}/survive"])         # bit-exact gate
    assert np.array_equal(pr["non_zero"], g[f"{n}/non_zero"])
    np.testing.assert_allclose(pr["pro_n"], g[f"{n}/pro_n"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("cfg_name", ["cfg1_vits_256", "cfg2_chaos_mri"])
def test_alp_config_shapes(cfg_name):
    g = _load("alp_configs.npz")
    seed, nq, L = [int(v) for v in g[f"{cfg_name}/meta"]]
    cfg = synth.CONFIGS[cfg_name]
    vol = synth.make_volume(seed, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None]
    for l in range(L):
        for q in range(nq):
            qry = np.transpose(vol.qry[q], (2, 0, 1))[None, None]
            for kind, mask, mode in (("bg", vol.bg[l], "gridconv"), ("fg", vol.fg[l], "gridconv+"),
                                     ("fgmask", vol.fg[l], "mask")):
                n = f"{cfg_name}/l{l}/q{q}/{kind}"
                pred, assign, _, grid = O.alp_forward(qry, sup_x, mask[None, :, None], mode, 0.95,
                                                      [cfg["h"] // 8, cfg["w"] // 8], isval=True, val_wsize=cfg["ws"])
                np.testing.assert_allclose(pred, g[f"{n}/pred_grid"], atol=MAP_TOL, rtol=0)
                assert np.array_equal(grid, g[f"{n}/proto_grid"])
                if mode != "mask":
                    pr = O.get_prototypes(np.transpose(vol.sup, (0, 3, 1, 2)), mask[:, None], mode,
                                          (cfg["ws"],) * 2, 0.95, cfg["ws"])
                    assert np.array_equal(pr["survive"], g[f"{n}/survive"])
                    mism = assign[0] != g[f"{n}/debug_assign"]
                    assert mism.mean() < 1e-3                     # argmax may flip only on fp32 near-ties


def _cfg2_keys():
    g = _load("alp_configs2.npz")
    return sorted({n.split("/")[0] for n in g["names"]})


@pytest.mark.parametrize("key", _cfg2_keys())
def test_alp_config_shapes_remaining(key):
    """BASELINE configs 3 and 4 and the window sweep of config 5 (ws 3, 6, 7) at C = 1024: reference maps, survival
    masks and assignments (alp_configs2.npz)."""
    g = _load("alp_configs2.npz")
    seed, nq, L, ws = [int(v) for v in g[f"{key}/meta"]]
    cfg = synth.CONFIGS[key if key in synth.CONFIGS else "cfg5_stress_vitl"]
    vol = synth.make_volume(seed, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None]
    qry = np.transpose(vol.qry[0], (2, 0, 1))[None, None]
    for n in [x for x in g["names"] if x.startswith(key + "/")]:
        _, lname, _, kind = n.split("/")
        l = int(lname[1:])
        mask = vol.bg[l] if kind == "bg" else vol.fg[l]
        mode = {"bg": "gridconv", "fg": "gridconv+", "fgmask": "mask"}[kind]
        pred, assign, _, _ = O.alp_forward(qry, sup_x, mask[None, :, None], mode, 0.95, [cfg["h"] // 8, cfg["w"] // 8],
                                           isval=True, val_wsize=ws)
        # fp32 summation order over C = 1024 channels and up to ~1300 prototypes differs between the oracle's loops and
        # ATen's conv: a handful of pixels sit at 1.2e-4 (north_star's bar is 1e-3)
        np.testing.assert_allclose(pred, g[f"{n}/pred_grid"], atol=2.5 * MAP_TOL, rtol=0)
        if mode != "mask":
            pr = O.get_prototypes(np.transpose(vol.sup, (0, 3, 1, 2)), mask[:, None], mode, (ws, ws), 0.95, ws)
            assert np.array_equal(pr["survive"], g[f"{n}/survive"])
            assert (assign[0] != g[f"{n}/debug_assign"]).mean() < 1e-3     # argmax may flip only on fp32 near-ties


def _prompt_names():
    return list(_load("prompts.npz")["names"])


@pytest.mark.parametrize("key", _prompt_names())
def test_prompts(key):
    """points, labels and boxes handed to SamPredictor.predict: bit-exact, dtype included."""
    g = _load("prompts.npz")
    name, cfg = key.split("/")
    use_cca, pm = cfg.startswith("cca1"), cfg.split("_", 1)[1]
    out = O.coarse_to_prompts(g[f"{name}/low"], int(g[f"{name}/S"]), 1024, use_cca=use_cca, point_mode=pm)
    assert np.array_equal(np.packbits(out["pred"]), g[f"{name}/pred_bits"])
    assert np.array_equal(out["p_fg"][::61, ::67], g[f"{name}/p_fg_sample"])
    ncalls = int(g[f"{key}/ncalls"])
    if out["empty"]:
        assert ncalls == 0
        return
    assert ncalls == len(out["bboxes"])
    assert out["points"].dtype == g[f"{key}/points"].dtype
    assert np.array_equal(out["points"], g[f"{key}/points"])
    assert np.array_equal(out["bboxes"], g[f"{key}/boxes"]) and out["bboxes"].dtype == np.int64
    assert np.all(g[f"{key}/multimask"] == (not use_cca))
    # ProtoSAM.predict_w_points_bbox labels every point 1 (models/ProtoSAM.py:508)
    assert np.all(g[f"{key}/point_labels"] == 1)


# ------------------------------------------------------------------------------ optional variants (variants.npz)

def _variant_names():
    return list(_load("variants.npz")["names"])


def _variant_setup(g, name, use_cca):
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    out = O.coarse_to_prompts(low, S, 1024, use_cca=use_cca, point_mode="both")
    _, p, _ = O.coarse_logits_to_probs(low, S, 1024)
    return out, p, (out["n"], out["labels"], out["stats"], out["centroids"])


@pytest.mark.parametrize("name", _variant_names())
@pytest.mark.parametrize("dev", [0, 1])
@pytest.mark.parametrize("use_cca", [False, True])
def test_negative_points(name, dev, use_cca):
    """ProtoSAM(use_neg_points=True): the points/labels every SamPredictor.predict call received, under both
    ``.cpu()`` semantics (models/ProtoSAM.py:361-434, 500-523)."""
    g = _load("variants.npz")
    key = f"{name}/neg_dev{dev}_cca{int(use_cca)}"
    out, p, cc = _variant_setup(g, name, use_cca)
    ncalls = int(g[f"{key}/ncalls"])
    if out["empty"]:
        assert ncalls == 0
        return
    neg = O.get_neg_points(cc, p, host_aliasing=(dev == 0))
    assert ncalls == len(neg) == len(out["points"])
    for i in range(ncalls):
        want_pts, want_lab = g[f"{key}/points"][i], g[f"{key}/point_labels"][i]
        n = int((want_lab >= 0).sum())
        rows = [out["points"][i]] + ([neg[i]] if neg[i] is not None else [])
        got = np.vstack(rows)
        assert got.shape == (n, 2) and np.array_equal(got, want_pts[:n])
        assert np.array_equal(want_lab[:n], [1] * len(out["points"][i]) + [0] * (n - len(out["points"][i])))


@pytest.mark.parametrize("name", _variant_names())
def test_mask_prompts(name):
    """ProtoSAM(use_mask=True): the mask_input arrays handed to SamPredictor.predict (models/ProtoSAM.py:452-498)."""
    g = _load("variants.npz")
    out, _, cc = _variant_setup(g, name, False)
    ncalls = int(g[f"{name}/mask/ncalls"])
    if out["empty"]:
        assert ncalls == 0
        return
    masks, ids = O.sam_mask_inputs(cc)
    assert masks.shape == (ncalls, 1, 256, 256) and str(masks.dtype) == str(g[f"{name}/mask/dtype"])
    assert set(np.unique(masks)) <= set(g[f"{name}/mask/values"].tolist()) | {10, 248}
    assert np.array_equal(np.unique(masks), g[f"{name}/mask/values"])
    assert np.array_equal(np.packbits(masks == 10), g[f"{name}/mask/fg_bits"])
    assert np.array_equal(ids, np.arange(1, ncalls + 1))


@pytest.mark.parametrize("name", _variant_names())
def test_coarse_pred_only_confidence(name):
    """ProtoSAM(coarse_pred_only=True) (models/ProtoSAM.py:580-590): confidence of util/utils.py:429-434, and with
    use_cca the (mask, confidence) of cca(..., return_conf=True)."""
    g = _load("variants.npz")
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    logits_S = O.upsample_bilinear(low, S)
    p = O.softmax2(logits_S)
    pred = (p[0, 1] > p[0, 0]).astype(np.uint8)
    assert np.array_equal(np.packbits(pred), g[f"{name}/coarse_cca0/pred_bits"])
    assert O.get_confidence_from_logits(logits_S) == pytest.approx(float(g[f"{name}/coarse_cca0/conf"]), rel=2e-6, abs=1e-7)
    kept, conf = O.cca(pred, p[0, 1], return_conf=True)
    assert np.array_equal(np.packbits(kept), g[f"{name}/coarse_cca1/pred_bits"])
    assert float(conf) == pytest.approx(float(g[f"{name}/coarse_cca1/conf"]), rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("name", _variant_names())
@pytest.mark.parametrize("use_cca", [False, True])
def test_protomedsam_boxes(name, use_cca):
    """ProtoMedSAM.forward (models/ProtoMedSAM.py:175-200): boxes handed to medsam_inference; the confidences are those of
    a softmax applied to probabilities (need_softmax, then util/utils.py:486 again)."""
    g = _load("variants.npz")
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    out = O.coarse_to_prompts_medsam(low, S, 1024, use_cca=use_cca)
    assert out["need_softmax"] == bool(g[f"{name}/medsam_need_softmax"])
    key = f"{name}/medsam_cca{int(use_cca)}"
    if out["empty"]:
        assert int(g[f"{key}/ncalls"]) == 0
        return
    assert out["boxes_1024"].dtype == g[f"{key}/boxes"].dtype and np.array_equal(out["boxes_1024"], g[f"{key}/boxes"])
    if not use_cca:
        want = g[f"{name}/medsam_conf"]
        got = np.array([float(out["conf"][k]) for k in sorted(out["conf"])])
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)


def _topk_names():
    return list(_load("topk.npz")["names"])


@pytest.mark.parametrize("name", _topk_names())
@pytest.mark.parametrize("k", [2, 5, 17])
def test_most_conf_points_any_k(name, k):
    """ProtoSAM.get_most_conf_points(output_p_fg, pred, k) for k > 1 (models/ProtoSAM.py:266-289), as the reference computes
    it on the 1024^2 probability map: locations in torch.topk's order (equal probabilities included) and the confidences."""
    g = _load("topk.npz")
    labels = g[f"{name}/k{k}/labels"]
    if len(labels) == 0:
        pytest.skip("no component with k pixels")
    out, p, cc = _variant_setup(g, name, False)
    for i, j in enumerate(labels):
        loc, conf = O.get_most_conf_points(p[0, 1], cc[1] == j, k)
        assert loc.dtype == np.int64 and np.array_equal(loc, g[f"{name}/k{k}/locations"][i]), (name, k, int(j))
        assert np.array_equal(np.array(conf, np.float64), g[f"{name}/k{k}/confidences"][i])
    with pytest.raises(RuntimeError):
        O.get_most_conf_points(p[0, 1], cc[1] == labels[0], int((cc[1] == labels[0]).sum()) + 1)
