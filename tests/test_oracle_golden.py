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
