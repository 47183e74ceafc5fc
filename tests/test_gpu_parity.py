"""Parity of the CUDA path (through the C ABI, ctypes) with the reference-generated golden fixtures
and with the CPU oracle.  Bars (BASELINE.json north_star): survival masks, component counts,
argmax points, boxes, centroids bit-exact; similarity / prediction maps within 1e-3 absolute."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from protosam_b200 import _lib, ops, prompts as PR, synth  # noqa: E402
from protosam_b200.alpmodule import MultiProtoAsConv  # noqa: E402
from protosam_b200.engine import CoarseVolumeEngine  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAP_TOL = 1e-3
DEV = "cuda:0"


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------------------ ALP module

@pytest.mark.parametrize("n", list(_load("alp_small.npz")["names"]))
def test_alp_module_vs_reference_golden(n, capsys):
    g = _load("alp_small.npz")
    mode, isval, vw, pg, thresh, q5d = g[f"{n}/meta"]
    isval, vw, pg, thresh = bool(int(isval)), (None if vw == "None" else int(vw)), int(pg), float(thresh)
    sup, qry, y = g[f"{n}/sup"], g[f"{n}/qry"], g[f"{n}/y"]
    S, h, w, C = sup.shape
    m = MultiProtoAsConv(proto_grid=[pg, pg], feature_hw=[h, w])
    assert len(list(m.parameters())) == 0 and len(m.state_dict()) == 0
    sup_x = _t(sup).permute(0, 3, 1, 2)[None, :, None]              # channels-last storage like the caller's
    q = _t(qry).permute(2, 0, 1)[None, None]
    if not int(q5d):
        q = q[:, 0]
    sup_y = _t(y)[None, :, None]
    with torch.no_grad():
        if f"{n}/error" in g.files:
            m.check_empty = True                     # the reference's behaviour: raise inside forward (one host sync)
            with pytest.raises(RuntimeError):
                m(q, sup_x, sup_y, mode, thresh, isval=isval, val_wsize=vw, vis_sim=True)
            assert "failed to find prototypes" in capsys.readouterr().out
            m.check_empty = False                    # default: no host sync in forward, NaN scores, the error on demand
            pred, _, _, _ = m(q, sup_x, sup_y, mode, thresh, isval=isval, val_wsize=vw)
            assert torch.isnan(pred).all()
            with pytest.raises(RuntimeError):
                m.raise_if_empty()
            return
        pred, assign, vis, grid = m(q, sup_x, sup_y, mode, thresh, isval=isval, val_wsize=vw, vis_sim=True)
    assert tuple(pred.shape) == g[f"{n}/pred_grid"].shape
    assert tuple(assign[0].shape) == g[f"{n}/debug_assign"].shape
    assert tuple(grid.shape) == g[f"{n}/proto_grid"].shape
    np.testing.assert_allclose(pred.cpu().numpy(), g[f"{n}/pred_grid"], atol=MAP_TOL, rtol=0)
    sims = vis["raw_local_sims"].cpu().numpy()
    assert sims.shape == g[f"{n}/raw_local_sims"].shape
    np.testing.assert_allclose(sims, g[f"{n}/raw_local_sims"], atol=MAP_TOL, rtol=0)
    assert np.array_equal(grid.cpu().numpy(), g[f"{n}/proto_grid"])
    assert vis["proto_assign"] is assign[0]
    if mode == "mask":
        np.testing.assert_allclose(assign[0].cpu().numpy(), g[f"{n}/debug_assign"], atol=MAP_TOL, rtol=0)
        return
    # argmax may differ only where two similarities tie within fp32 noise
    a, ra = assign[0].cpu().numpy(), g[f"{n}/debug_assign"]
    bad = np.argwhere(a != ra)
    for (b, yy, xx) in bad:
        s = g[f"{n}/raw_local_sims"][b, :, yy, xx]
        assert abs(s[int(a[b, yy, xx])] - s[int(ra[b, yy, xx])]) < 1e-4
    # survival mask: bit-exact gate
    ks = (vw, vw) if isval else (h // pg, w // pg)
    pr = ops.alp_prototypes(_t(sup).permute(0, 3, 1, 2), _t(y)[None], [mode], ks, thresh)
    N = pr["N"]
    assert np.array_equal(pr["survive"][0, :N].cpu().numpy().astype(bool), g[f"{n}/survive"])
    P = int(pr["counts"][0])
    assert P == g[f"{n}/pro_n"].shape[0]
    np.testing.assert_allclose(pr["protos"][0, :P].cpu().numpy(), g[f"{n}/pro_n"], atol=1e-5, rtol=0)


def test_alp_module_rejects_cpu_and_grad():
    m = MultiProtoAsConv([8, 8], [16, 16])
    q = torch.zeros(1, 1, 8, 16, 16)
    sx = torch.zeros(1, 1, 1, 8, 16, 16)
    sy = torch.ones(1, 1, 1, 16, 16)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(q, sx, sy, "gridconv+", 0.95)                           # CPU tensors: no fallback
    with pytest.raises(ValueError):
        m(q, sx, sy, "bogus", 0.95)
    with pytest.raises(RuntimeError):
        m(q.to(DEV).requires_grad_(True), sx.to(DEV), sy.to(DEV), "gridconv+", 0.95)


def test_alp_contiguous_nchw_query_and_batched_queries():
    """NCHW-contiguous inputs (not channels-last) and several query slices in one call."""
    vol = synth.make_volume(3, Q=3, L=1, C=64, h=16, w=16, img_size=128)
    m = MultiProtoAsConv([8, 8], [16, 16])
    sx = _t(np.transpose(vol.sup, (0, 3, 1, 2)))[None, :, None]     # truly NCHW contiguous
    q = _t(np.transpose(vol.qry, (0, 3, 1, 2)))                      # [3,C,h,w]
    with torch.no_grad():
        pred, _, _, _ = m(q, sx, _t(vol.fg[0])[None, :, None], "gridconv+", 0.95, isval=True, val_wsize=2)
    for i in range(3):
        ref, _, _, _ = O.alp_forward(np.transpose(vol.qry[i], (2, 0, 1))[None], np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None],
                                     vol.fg[0][None, :, None], "gridconv+", 0.95, [2, 2], isval=True, val_wsize=2)
        np.testing.assert_allclose(pred[i].cpu().numpy(), ref[0], atol=MAP_TOL, rtol=0)


@pytest.mark.parametrize("cfg_name", ["cfg1_vits_256", "cfg2_chaos_mri"])
def test_engine_match_vs_reference_golden(cfg_name):
    """Batched engine (all labels, all slices, bg+fg sets in one launch) vs per-call reference outputs."""
    g = _load("alp_configs.npz")
    seed, nq, L = [int(v) for v in g[f"{cfg_name}/meta"]]
    cfg = synth.CONFIGS[cfg_name]
    vol = synth.make_volume(seed, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    for fg_mode, kind in (("gridconv+", "fg"), ("mask", "fgmask")):
        eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"], fg_mode=fg_mode)
        pr = eng.set_support(_t(vol.sup), _t(vol.fg))
        logits = eng.match(_t(vol.qry)).cpu().numpy()               # [Q*L,2,h,w]
        N = pr["N"]
        for l in range(L):
            for name, si in (("bg", 2 * l), (kind, 2 * l + 1)):
                if name != "fgmask":
                    assert np.array_equal(pr["survive"][si, :N].cpu().numpy().astype(bool), g[f"{cfg_name}/l{l}/q0/{name}/survive"])
                for q in range(nq):
                    ref = g[f"{cfg_name}/l{l}/q{q}/{name}/pred_grid"][0, 0]
                    np.testing.assert_allclose(logits[q * L + l, si % 2], ref, atol=MAP_TOL, rtol=0)


@pytest.mark.parametrize("key", ["cfg3_synapse_ct", "cfg4_polyp_1024", "cfg5_ws3", "cfg5_ws6", "cfg5_ws7"])
def test_engine_match_vs_reference_golden_remaining_configs(key):
    """BASELINE configs 3, 4 and the ws = 3, 6, 7 points of config 5 (C = 1024) against maps and survival masks the
    unmodified reference produced (tests/golden/alp_configs2.npz)."""
    g = _load("alp_configs2.npz")
    seed, nq, L, ws = [int(v) for v in g[f"{key}/meta"]]
    cfg = synth.CONFIGS[key if key in synth.CONFIGS else "cfg5_stress_vitl"]
    vol = synth.make_volume(seed, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    kinds = sorted({n.split("/")[3] for n in g["names"] if n.startswith(key + "/")})
    for fg_mode, kind in (("gridconv+", "fg"), ("mask", "fgmask")):
        if kind not in kinds:
            continue
        eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=ws, fg_mode=fg_mode)
        pr = eng.set_support(_t(vol.sup), _t(vol.fg))
        logits = eng.match(_t(vol.qry)).cpu().numpy()               # [Q*L,2,h,w]
        N = pr["N"]
        for l in range(L):
            for name, si in (("bg", 2 * l), (kind, 2 * l + 1)):
                if name != "fgmask":
                    assert np.array_equal(pr["survive"][si, :N].cpu().numpy().astype(bool), g[f"{key}/l{l}/q0/{name}/survive"])
                ref = g[f"{key}/l{l}/q0/{name}/pred_grid"][0, 0]
                np.testing.assert_allclose(logits[l, si % 2], ref, atol=MAP_TOL, rtol=0)


def test_engine_auto_fg_mode_matches_caller_rule():
    """grid_proto_fewshot.py:254-256: 'gridconv+' iff some kernel_size window is >= 0.95 foreground."""
    h = w = 32
    fg = np.zeros((3, 1, h, w), np.float32)
    fg[0, 0, 4:8, 4:8] = 1          # exactly one full 4x4 training window -> gridconv+
    fg[1, 0, 5:9, 5:9] = 1          # 4x4 blob straddling windows -> no full window -> mask
    fg[2, 0, 10, 10] = 1            # single pixel -> mask
    sup = synth.layer_norm(synth.gaussian_like(1, (1, h, w, 32)))
    eng = CoarseVolumeEngine((h, w), 256, val_wsize=2, proto_grid_size=8)
    pr = eng.set_support(_t(sup), _t(fg))
    eff = pr["eff_modes"].cpu().numpy()
    assert [_lib.MODE_NAMES[int(e)] for e in eff] == ["gridconv", "gridconv+", "gridconv", "mask", "gridconv", "mask"]
    counts = pr["counts"].cpu().numpy()
    assert counts[1] == 4 + 1 and counts[3] == 1 and counts[5] == 1


# ------------------------------------------------------------------------------ prompts

def _prompt_cases():
    return list(_load("prompts.npz")["names"])


@pytest.mark.parametrize("key", _prompt_cases())
def test_prompts_vs_reference_golden(key):
    """What ProtoSAM.forward hands to SamPredictor.predict, bit for bit (dtype included)."""
    g = _load("prompts.npz")
    name, cfg = key.split("/")
    use_cca, pm = cfg.startswith("cca1"), cfg.split("_", 1)[1]
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    sp = PR.coarse_to_prompts(_t(low), S, 1024, use_cca=use_cca, point_mode=pm, max_cc=512)[0]
    calls = sp.predict_calls()
    assert len(calls) == int(g[f"{key}/ncalls"])
    if not calls:
        return
    pts = np.stack([c["point_coords"] for c in calls])
    assert pts.dtype == g[f"{key}/points"].dtype and np.array_equal(pts, g[f"{key}/points"])
    assert np.array_equal(np.stack([c["box"] for c in calls]), g[f"{key}/boxes"])
    assert np.array_equal(np.stack([c["point_labels"] for c in calls]), g[f"{key}/point_labels"])
    assert all(c["multimask_output"] == bool(m) for c, m in zip(calls, g[f"{key}/multimask"]))


@pytest.mark.parametrize("name", sorted({k.split("/")[0] for k in _prompt_cases()}))
def test_upsample_softmax_bits_vs_reference_golden(name):
    g = _load("prompts.npz")
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    p_fg, bits, probs2 = ops.upsample_softmax(_t(low), S, 1024, want_probs2=True)
    mask = np.unpackbits(bits[0].cpu().numpy().view(np.uint8), bitorder="little").reshape(1024, 1024)
    assert np.array_equal(np.packbits(mask), g[f"{name}/pred_bits"])
    assert np.array_equal(p_fg[0].cpu().numpy()[::61, ::67], g[f"{name}/p_fg_sample"])
    assert torch.equal(probs2[0, 1], p_fg[0])


@pytest.mark.parametrize("h,S,out", [(32, 256, 1024), (37, 518, 1024), (48, 672, 1024), (73, 1024, 1024),
                                     (24, 96, 256), (16, 64, 64)])
def test_upsample_softmax_full_map_vs_oracle(h, S, out):
    """every pixel of p (both channels) equals the oracle's ATen-CPU restatement"""
    low = (synth.gaussian_like(h * 7 + S, (3, 2, h, h)) * 9).astype(np.float32)
    p_fg, bits, probs2 = ops.upsample_softmax(_t(low), S, out, want_probs2=True)
    for i in range(3):
        _, p, pred = O.coarse_logits_to_probs(low[i:i + 1], S, out)
        assert np.array_equal(probs2[i].cpu().numpy(), p[0])
        mask = np.unpackbits(bits[i].cpu().numpy().view(np.uint8), bitorder="little").reshape(out, out)
        assert np.array_equal(mask, pred)


def test_full_resolution_logits_path():
    lg = (synth.gaussian_like(11, (1, 2, 256, 256)) * 6).astype(np.float32)
    p_fg, bits, probs2 = ops.upsample_softmax(_t(lg), 256, 256, want_probs2=True)
    assert np.array_equal(probs2.cpu().numpy(), O.softmax2(lg))


def _pack(mask):
    return torch.from_numpy(np.packbits(mask.astype(np.uint8), axis=-1, bitorder="little").view(np.int32)).to(DEV)


@pytest.mark.parametrize("out,density,seed", [(64, 0.5, 0), (256, 0.4, 1), (1024, 0.45, 2), (1024, 0.6, 3),
                                              (1024, 0.02, 4), (1024, 0.98, 5)])
def test_components_vs_oracle_noise_masks(out, density, seed):
    """pixel-noise masks (up to ~70k components): labels, order, stats, centroids, points, confidences"""
    rng = np.random.default_rng(seed)
    mask = rng.random((out, out)) < density
    p = (0.5 + rng.integers(0, 1 << 23, (out, out)) * 2.0 ** -24).astype(np.float32)
    p[rng.random((out, out)) < 0.3] = 1.0                              # many ties at the maximum
    n, lab, st, ce = O.connected_components(mask)
    hdr, recs, labels = ops.components(_pack(mask)[None], _t(p)[None], use_cca=False, max_cc=1 << 17,
                                       max_runs=out * out // 2 + 1, want_labels=True)
    H, R = ops.decode_headers(hdr)[0], ops.decode_records(recs)[0]
    assert int(H["flags"]) == 0 and int(H["ncc"]) == n - 1 and int(H["n_fg"]) == int(mask.sum())
    assert np.array_equal(labels[0].cpu().numpy(), lab)
    R = R[: n - 1]
    assert np.array_equal(R["label"], np.arange(1, n))
    assert np.array_equal(R["area"], st[1:, 4])
    assert np.array_equal(R["box"][:, 0], st[1:, 0]) and np.array_equal(R["box"][:, 1], st[1:, 1])
    assert np.array_equal(R["box"][:, 2] - R["box"][:, 0] + 1, st[1:, 2])
    assert np.array_equal(R["centroid"], ce[1:])
    assert np.array_equal(H["bg_stats"], st[0]) and np.array_equal(H["bg_centroid"], ce[0])
    bbox, pt, val = O._cc_prompts(p, lab, n)
    assert np.array_equal(R["box"], bbox[1:])
    assert np.array_equal(R["conf_pt"], pt[1:])                        # incl. torch.topk's n<64 tie rule
    assert np.array_equal(R["conf_pt_p"], val[1:])
    sums = O.cc_sums(p, lab, n).astype(np.float64)[1:] / (mask.sum() + 1e-6)
    np.testing.assert_allclose(R["conf"], sums, rtol=1e-5)


def test_components_use_cca_selects_oracle_component():
    rng = np.random.default_rng(9)
    for t in range(4):
        low = (rng.random((20, 20)) < 0.25).astype(np.float32)
        up = torch.nn.functional.interpolate(torch.from_numpy(low)[None, None], size=(512, 512), mode="bilinear")[0, 0].numpy()
        mask = up > 0.35
        p = (0.5 + np.minimum(up, 1.0) * 0.4999).astype(np.float32)
        cc = O.cca(mask.astype(np.uint8), p, return_cc=True)
        hdr, recs, labels = ops.components(_pack(mask)[None], _t(p)[None], use_cca=True, max_cc=1, want_labels=True)
        H, R = ops.decode_headers(hdr)[0], ops.decode_records(recs)[0]
        assert int(H["n_rec"]) == 1 and np.array_equal(labels[0].cpu().numpy(), cc[1])
        assert np.array_equal(R[0]["centroid"], cc[3][1]) and int(R[0]["area"]) == cc[2][1, 4]


def test_run_table_overflow_is_reported():
    mask = np.zeros((256, 256), bool)
    mask[:, ::2] = True
    hdr, _, _ = ops.components(_pack(mask)[None], _t(np.full((256, 256), 0.75, np.float32))[None], max_runs=1000)
    H = ops.decode_headers(hdr)[0]
    assert int(H["flags"]) & _lib.IMG_RUN_OVERFLOW
    with pytest.raises(RuntimeError):
        PR.prompts_from_records(H, None, False)


def test_function_level_dropins_match_oracle():
    lg = (torch.nn.functional.interpolate(torch.from_numpy(synth.gaussian_like(21, (1, 2, 9, 9)) * 9), size=(512, 512),
                                          mode="bicubic")).numpy().astype(np.float32)
    p = O.softmax2(lg)
    pred = (p[0, 1] > p[0, 0]).astype(np.uint8)
    cc, conf = PR.get_connected_components(pred, _t(lg), return_conf=True)
    rcc, rconf = O.get_connected_components(pred, p[0, 1], return_conf=True)
    assert cc[0] == rcc[0] and np.array_equal(cc[1], rcc[1]) and np.array_equal(cc[2], rcc[2]) and np.array_equal(cc[3], rcc[3])
    for k in rconf:
        assert conf[k] == pytest.approx(float(rconf[k]), rel=1e-5)
    assert np.array_equal(PR.get_bbox_per_cc(cc), O.get_bbox_per_cc(rcc))
    for pm in ("conf", "centroid", "both"):
        pts, labels, _, _ = PR.get_sam_input_points(cc, None, point_mode=pm)
        rpts, rlabels = O.get_sam_input_points(rcc, p[0, 1], pm)
        assert pts.dtype == rpts.dtype and np.array_equal(pts, rpts) and np.array_equal(labels, rlabels)
    # get_most_conf_points per component == torch.topk(p_fg[labels == j], 1) of the reference
    import torch as _torch
    for j in range(1, cc[0]):
        loc, conf_j = PR.get_most_conf_points(cc, j, 1)
        m = _torch.from_numpy(rcc[1] == j)
        v, i = _torch.topk(_torch.from_numpy(p[0, 1])[m], 1)
        ref_loc = _torch.nonzero(m)[i][:, [1, 0]].numpy()
        assert np.array_equal(loc, ref_loc) and conf_j[0] == float(v[0])
    sel = PR.cca(pred, _t(lg), return_cc=True)
    rsel = O.cca(pred, p[0, 1], return_cc=True)
    assert sel[0] == rsel[0] and np.array_equal(sel[1], rsel[1]) and np.array_equal(sel[3], rsel[3])
    assert np.array_equal(PR.cca(pred, _t(lg)), O.cca(pred, p[0, 1]))


@pytest.mark.parametrize("scale", [9.0, 40.0])
def test_most_conf_points_any_k_matches_torch_topk(scale):
    """get_most_conf_points(output_p_fg, pred, k) for k > 1 (models/ProtoSAM.py:266-289): the device replay of torch.topk
    == torch.topk itself on the reference's masked map, for components on both sides of ATen's k * 64 <= n switch and
    with saturated (exactly equal) probabilities at scale 40."""
    import torch as _torch
    lg = (torch.nn.functional.interpolate(torch.from_numpy(synth.gaussian_like(33, (1, 2, 11, 11)) * scale), size=(384, 384),
                                          mode="bicubic")).numpy().astype(np.float32)
    # speckle: many small components (nth_element + sort side) next to the big ones (partial_sort side)
    rng = np.random.default_rng(3)
    lg[0, 1] += (rng.random((384, 384)) < 0.02).astype(np.float32) * 2 * scale
    p = O.softmax2(lg)
    pred = (p[0, 1] > p[0, 0]).astype(np.uint8)
    cc, _ = PR.get_connected_components(pred, _t(lg), return_conf=True)
    rcc, _ = O.get_connected_components(pred, p[0, 1], return_conf=True)
    assert cc[0] == rcc[0] and cc[0] > 20
    areas = cc[2][:, 4]
    pf = _torch.from_numpy(p[0, 1])
    checked = {"partial_sort": 0, "nth_element": 0, "raises": 0}
    for k in (2, 3, 5, 17):
        for j in range(1, cc[0]):
            if areas[j] < k:
                with pytest.raises(RuntimeError):
                    PR.get_most_conf_points(cc, j, k)
                checked["raises"] += 1
                continue
            if areas[j] > 2000 and j % 3:            # keep the runtime down: every third big component
                continue
            m = _torch.from_numpy(rcc[1] == j)
            v, i = _torch.topk(pf[m], k)
            ref_loc = _torch.nonzero(m)[i][:, [1, 0]].numpy()
            loc, conf = PR.get_most_conf_points(cc, j, k)
            assert loc.dtype == ref_loc.dtype and np.array_equal(loc, ref_loc), (k, j, int(areas[j]))
            assert conf == [float(x) for x in v]
            checked["partial_sort" if k * 64 <= areas[j] else "nth_element"] += 1
    assert checked["partial_sort"] > 0 and checked["nth_element"] > 10


# ------------------------------------------------------------------------------ whole path

@pytest.mark.parametrize("cfg_name,nq", [("cfg1_vits_256", 1), ("cfg2_chaos_mri", 3), ("cfg3_synapse_ct", 2)])
@pytest.mark.parametrize("use_cca", [False, True])
def test_engine_end_to_end_vs_oracle(cfg_name, nq, use_cca):
    """features -> prompts on the device vs the oracle run stage by stage on the same inputs."""
    cfg = synth.CONFIGS[cfg_name]
    L = min(cfg["L"], 2)
    vol = synth.make_volume(4321, Q=nq, L=L, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"], use_cca=use_cca)
    eng.set_support(_t(vol.sup), _t(vol.fg))
    q = _t(vol.qry)
    logits = eng.match(q).cpu().numpy()
    got = eng.decode(*eng.run(q))
    for qi in range(nq):
        for l in range(L):
            ref = O.coarse_to_prompts(logits[qi * L + l][None], cfg["img_size"], 1024, use_cca=use_cca, point_mode="both")
            s = got[qi][l]
            assert s.empty == ref["empty"]
            if s.empty:
                continue
            assert np.array_equal(s.boxes, ref["bboxes"])
            assert np.array_equal(s.points, ref["points"]) and s.points.dtype == ref["points"].dtype
            assert np.array_equal(s.point_labels, ref["point_labels"])


def test_full_size_properties_cfg2():
    """BASELINE config 2 at full size (32 slices x 4 labels): size-independent properties."""
    cfg = synth.CONFIGS["cfg2_chaos_mri"]
    vol = synth.make_volume(99, Q=cfg["Q"], L=cfg["L"], C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"])
    eng.set_support(_t(vol.sup), _t(vol.fg))
    q = _t(vol.qry)
    hdr, recs = eng.run(q)
    H, R = ops.decode_headers(hdr), ops.decode_records(recs)
    assert len(H) == cfg["Q"] * cfg["L"] and not (H["flags"] & (_lib.IMG_RUN_OVERFLOW | _lib.IMG_CC_TRUNCATED)).any()
    for i in range(len(H)):
        r = R[i, : H["n_rec"][i]]
        assert r["area"].sum() == H["n_fg"][i]                          # components partition the foreground
        assert (r["box"][:, 0] <= r["conf_pt"][:, 0]).all() and (r["conf_pt"][:, 0] <= r["box"][:, 2]).all()
        assert (r["box"][:, 1] <= r["conf_pt"][:, 1]).all() and (r["conf_pt"][:, 1] <= r["box"][:, 3]).all()
        assert np.all(np.diff(r["label"]) == 1)
    # determinism + slice-permutation equivariance (slices are independent given the prototypes)
    hdr2, recs2 = eng.run(q)
    assert torch.equal(hdr, hdr2) and torch.equal(recs, recs2)
    perm = torch.randperm(cfg["Q"], generator=torch.Generator().manual_seed(0)).to(DEV)
    hdr3, recs3 = eng.run(q[perm].contiguous())
    Lb = cfg["L"]
    idx = (perm[:, None] * Lb + torch.arange(Lb, device=DEV)[None]).reshape(-1)
    assert torch.equal(hdr3, hdr[idx]) and torch.equal(recs3, recs[idx])


@pytest.mark.parametrize("h,S", [(37, 518), (73, 1024), (48, 672)])
def test_engine_variant_of_upsample_equals_full_variant(h, S):
    """fg-only evaluation + per-word statistics give the same mask, probabilities and records as
    the every-pixel variant followed by pixel-wise statistics."""
    low = np.concatenate([(synth.gaussian_like(h + S, (2, 2, h, h)) * 9), synth.gaussian_like(5, (1, 2, h, h)) * 1e-4]).astype(np.float32)
    p_full, bits_full, _ = ops.upsample_softmax(_t(low), S, 1024)
    p_fg, bits, _, wstat = ops.upsample_softmax(_t(low), S, 1024, fg_only=True, want_wstat=True)
    assert torch.equal(bits, bits_full)
    mask = torch.from_numpy(np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little").reshape(3, 1024, 1024)).bool().to(DEV)
    # the engine variant writes p_fg only where kernel 3b can read it per pixel: always in words that are not full,
    # in full words only next to a non-full word above/below (or at a block's first/last row); what it writes is exact
    partial = (bits != -1).repeat_interleave(32, dim=2) & mask
    assert partial.any() and torch.equal(p_fg[partial], p_full[partial])
    written = mask & (p_fg != 0)
    assert torch.equal(p_fg[written], p_full[written])
    full_w = (bits == -1)
    lonely = full_w.clone()
    lonely[:, 1:-1] &= ~(full_w[:, :-2] & full_w[:, 2:])          # full words without two full vertical neighbours
    lonely_px = lonely.repeat_interleave(32, dim=2)
    assert torch.equal(p_fg[lonely_px], p_full[lonely_px])
    for use_cca in (False, True):
        a = ops.components(bits, p_fg, use_cca=use_cca, max_cc=4096, wstat=wstat)
        b = ops.components(bits_full, p_full, use_cca=use_cca, max_cc=4096)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ------------------------------------------------------------------------------ tensor-core match

def _fake_protos(rng, counts, modes, cap, C):
    """a prototype table as psam_alp_prototypes would write it: unit rows, arbitrary per-set counts"""
    nsets = len(counts)
    pr = rng.standard_normal((nsets, cap, C)).astype(np.float32)
    pr /= np.linalg.norm(pr, axis=-1, keepdims=True)
    return dict(protos=_t(pr), counts=_t(np.array(counts, np.int32)), cap_rows=cap,
                eff_modes=_t(np.array([_lib.MODE_IDS[m] for m in modes], np.int32)),
                status=torch.zeros(nsets, dtype=torch.int32, device=DEV))


@pytest.mark.parametrize("Q,HW,C,counts,modes", [
    (1, 128, 64, [16], ["gridconv"]),                                            # one tile, one k-block, one group
    (1, 100, 64, [5], ["gridconv+"]),                                            # partial tile, partial group
    (2, 1369, 768, [300, 9, 287, 1, 310, 33, 256, 17], ["gridconv", "gridconv+", "gridconv", "mask"] * 2),
    (3, 1024, 384, [255, 1], ["gridconv", "mask"]),
    (1, 2304, 1024, [577, 40], ["gridconv", "gridconv+"]),                      # a set spanning three chunks
    (2, 333, 72, [0, 20, 0, 3], ["gridconv", "gridconv+", "gridconv", "mask"]),  # empty sets, C not a multiple of 64
    (5, 200, 128, [1] * 40, ["mask"] * 40),                                      # many tiny sets
    (40, 1369, 256, [100, 7], ["gridconv", "gridconv+"]),                        # more tiles than SMs
])
def test_match_tensor_core_vs_cuda_core(Q, HW, C, counts, modes):
    """the tcgen05 split-bf16 kernel agrees with the exact-fp32 kernel far inside the 1e-3 budget"""
    rng = np.random.default_rng(Q * 1000 + HW + C)
    cap = max(max(counts), 1) + 3
    pr = _fake_protos(rng, counts, modes, cap, C)
    qry = _t((rng.standard_normal((Q, HW, C)) * rng.uniform(0.1, 5.0, (Q, HW, 1))).astype(np.float32))
    # make some queries close to a prototype so that the softmax is peaked, and one all-zero row
    qry[0, : min(HW, 50)] += 4.0 * pr["protos"][-1, 0]
    qry[-1, HW - 1] = 0.0
    s1, a1, _ = ops.alp_match(qry, pr, want_assign=True, algo=1)
    st1 = pr["status"].clone()
    pr["status"].zero_()
    s2, a2, _ = ops.alp_match(qry, pr, want_assign=True, algo=2)
    torch.cuda.synchronize()
    assert torch.equal(st1, pr["status"])
    s1, s2, a1, a2 = s1.cpu().numpy(), s2.cpu().numpy(), a1.cpu().numpy(), a2.cpu().numpy()
    for i, (c, m) in enumerate(zip(counts, modes)):
        if c == 0:
            assert np.isnan(s2[:, i]).all() and np.isnan(a2[:, i]).all()
            continue
        np.testing.assert_allclose(s2[:, i], s1[:, i], atol=1e-4, rtol=0)
        if m == "mask":
            np.testing.assert_allclose(a2[:, i], a1[:, i], atol=1e-4, rtol=0)
        else:
            # argmax may differ only between near-tied prototypes
            bad = np.argwhere(a1[:, i] != a2[:, i])
            assert len(bad) <= 0.001 * a1[:, i].size + 2
    # the fused variant (query converted inside the GEMM) computes the same products in the same order; only the
    # row norm is summed in a different order
    pr["status"].zero_()
    s3, a3, _ = ops.alp_match(qry, pr, want_assign=True, algo=3)
    torch.cuda.synchronize()
    assert torch.equal(st1, pr["status"])
    s3, a3 = s3.cpu().numpy(), a3.cpu().numpy()
    for i, (c, m) in enumerate(zip(counts, modes)):
        if c == 0:
            assert np.isnan(s3[:, i]).all() and np.isnan(a3[:, i]).all()
        else:
            np.testing.assert_allclose(s3[:, i], s2[:, i], atol=2e-5, rtol=0)
            if m != "mask":
                assert (a3[:, i] != a2[:, i]).sum() <= 0.001 * a2[:, i].size + 2
    # auto (algo 0) takes a tensor-core path for these shapes
    s0, _, _ = ops.alp_match(qry, pr, want_assign=False, algo=0)
    s0 = s0.cpu().numpy()
    assert np.array_equal(s0, s2, equal_nan=True) or np.array_equal(s0, s3, equal_nan=True)


def test_match_auto_falls_back_to_cuda_cores_for_sims_and_odd_channels():
    rng = np.random.default_rng(5)
    pr = _fake_protos(rng, [10, 3], ["gridconv", "gridconv+"], 12, 20)      # C = 20: not a multiple of 8
    qry = _t(rng.standard_normal((1, 50, 20)).astype(np.float32))
    s0, _, _ = ops.alp_match(qry, pr, algo=0)
    s1, _, _ = ops.alp_match(qry, pr, algo=1)
    assert torch.equal(s0, s1)
    with pytest.raises(RuntimeError):
        ops.alp_match(qry, pr, algo=2)
    pr = _fake_protos(rng, [10, 3], ["gridconv", "gridconv+"], 12, 64)
    qry = _t(rng.standard_normal((1, 50, 64)).astype(np.float32))
    s0, _, sims0 = ops.alp_match(qry, pr, want_sims=True, algo=0)
    s1, _, sims1 = ops.alp_match(qry, pr, want_sims=True, algo=1)
    assert torch.equal(s0, s1) and torch.equal(sims0[:, 0, :10], sims1[:, 0, :10])


# ------------------------------------------------------------------------------ remaining BASELINE configs

def _oracle_logits(vol, cfg, ws, q, l, fg_mode, ks):
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None]
    qry = np.transpose(vol.qry[q], (2, 0, 1))[None]
    bg, _, _, _ = O.alp_forward(qry, sup_x, vol.bg[l][None, :, None], "gridconv", 0.95, ks, isval=True, val_wsize=ws)
    fg, _, _, _ = O.alp_forward(qry, sup_x, vol.fg[l][None, :, None], fg_mode, 0.95, ks, isval=True, val_wsize=ws)
    return np.concatenate([bg, fg], 1)[0]


@pytest.mark.parametrize("ws", [2, 3, 4, 5, 6, 7, 8])
def test_cfg5_window_sweep_vs_oracle(ws):
    """BASELINE config 5 (ViT-L/14 1024-d, 48x48, grid window 2..8, 'gridconv+') at 2 slices: maps within 1e-3 of the
    oracle, prompts bit-exact on the GPU's own maps."""
    cfg = synth.CONFIGS["cfg5_stress_vitl"]
    vol = synth.make_volume(50 + ws, Q=2, L=1, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=ws, fg_mode="gridconv+")
    pr = eng.set_support(_t(vol.sup), _t(vol.fg))
    q = _t(vol.qry)
    logits = eng.match(q).cpu().numpy()
    ks = [cfg["h"] // 8, cfg["w"] // 8]
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))
    for name, mask, mode, si in (("bg", vol.bg[0], "gridconv", 0), ("fg", vol.fg[0], "gridconv+", 1)):
        ref = O.get_prototypes(sup_x, mask[:, None], mode, (ws, ws), 0.95)
        assert np.array_equal(ref["survive"], pr["survive"][si, : pr["N"]].cpu().numpy().astype(bool)), name
    got = eng.decode(*eng.run(q))
    for qi in range(2):
        ref = _oracle_logits(vol, cfg, ws, qi, 0, "gridconv+", ks)
        np.testing.assert_allclose(logits[qi], ref, atol=MAP_TOL, rtol=0)
        rp = O.coarse_to_prompts(logits[qi][None], cfg["img_size"], 1024, use_cca=False, point_mode="both")
        s = got[qi][0]
        assert s.empty == rp["empty"]
        if not s.empty:
            assert np.array_equal(s.boxes, rp["bboxes"]) and np.array_equal(s.points, rp["points"])


@pytest.mark.parametrize("hw", [48, 73])
@pytest.mark.parametrize("point_mode", ["conf", "both"])
def test_cfg4_polyp_mask_and_gridconv_single_stage(hw, point_mode):
    """BASELINE config 4: fg forced to 'mask' (small lesion), bg 'gridconv', logits upsampled straight to 1024."""
    cfg = synth.CONFIGS["cfg4_polyp_1024"]
    vol = synth.make_volume(404 + hw, Q=2, L=1, C=cfg["C"], h=hw, w=hw, img_size=1024)
    eng = CoarseVolumeEngine((hw, hw), 1024, out_size=1024, val_wsize=2, fg_mode="mask", point_mode=point_mode)
    eng.set_support(_t(vol.sup), _t(vol.fg))
    q = _t(vol.qry)
    logits = eng.match(q).cpu().numpy()
    got = eng.decode(*eng.run(q))
    cfg2 = dict(cfg, h=hw, w=hw)
    for qi in range(2):
        ref = _oracle_logits(vol, cfg2, 2, qi, 0, "mask", [hw // 8, hw // 8])
        np.testing.assert_allclose(logits[qi], ref, atol=MAP_TOL, rtol=0)
        rp = O.coarse_to_prompts(logits[qi][None], 1024, 1024, use_cca=False, point_mode=point_mode)
        s = got[qi][0]
        assert s.empty == rp["empty"]
        if not s.empty:
            assert np.array_equal(s.boxes, rp["bboxes"])
            assert np.array_equal(s.points, rp["points"]) and s.points.dtype == rp["points"].dtype


def test_mask_nearest_matches_aten_and_feeds_the_engine():
    """the caller-side nearest resize of the support masks (grid_proto_fewshot.py:228-231) on the device"""
    for (H, hh) in ((256, 32), (518, 37), (672, 48), (1024, 73), (37, 37)):
        m = (torch.rand((2, 1, H, H), generator=torch.Generator().manual_seed(H)) > 0.6).float()
        ref = torch.nn.functional.interpolate(m, size=(hh, hh), mode="nearest")
        got = ops.mask_nearest(m.to(DEV), hh, hh)
        assert torch.equal(got.cpu(), ref)
    cfg = synth.CONFIGS["cfg2_chaos_mri"]
    vol = synth.make_volume(77, Q=1, L=2, C=64, h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"])
    pa = eng.set_support_from_image_masks(_t(vol.sup), _t(vol.fg_img))
    a, ca = pa["protos"].clone(), pa["counts"].clone()
    pb = eng.set_support(_t(vol.sup), _t(vol.fg))
    assert torch.equal(ca, pb["counts"])
    for i, c in enumerate(ca.tolist()):          # rows beyond a set's count are unspecified
        assert torch.equal(a[i, :c], pb["protos"][i, :c])


# ------------------------------------------------------------------------------ degenerate inputs

def test_engine_degenerate_masks_and_queries_vs_oracle():
    """empty foreground, full foreground (empty background set), a mask smaller than any window, an all-zero query
    pixel and an all-zero query slice: status bits, modes and maps agree with the oracle / the reference's rules"""
    h = w = 16
    C, img = 64, 128
    vol = synth.make_volume(31, Q=3, L=1, C=C, h=h, w=w, img_size=img)
    fg = np.zeros((4, 1, h, w), np.float32)
    fg[1] = 1.0                       # label 1: everything foreground -> bg 'gridconv' set is empty
    fg[2, 0, 3, 5] = 1.0              # label 2: one pixel -> fg decided 'mask', no local window survives
    fg[3, 0, 4:12, 4:12] = 1.0        # label 3: a regular blob
    qry = vol.qry.copy()
    qry[1] = 0.0                      # all-zero slice: norms clamp at 1e-4, every similarity is exactly 0
    qry[2, 0, 0] = 0.0                # one all-zero pixel
    eng = CoarseVolumeEngine((h, w), img, out_size=256, val_wsize=2, proto_grid_size=8)
    pr = eng.set_support(_t(vol.sup), _t(fg))
    logits = eng.match(_t(qry)).cpu().numpy().reshape(3, 4, 2, h, w)
    torch.cuda.synchronize()
    status, eff, counts = pr["status"].cpu().numpy(), pr["eff_modes"].cpu().numpy(), pr["counts"].cpu().numpy()
    names = [_lib.MODE_NAMES[int(e)] for e in eff]
    # label 0: no foreground at all -> fg 'mask' with a zero prototype; bg uses every window
    assert names[0] == "gridconv" and names[1] == "mask" and counts[0] == 64 and counts[1] == 1
    # label 1: bg set empty -> flagged, NaN scores (the reference raises inside F.conv2d, alpmodule.py:68)
    assert status[2] & _lib.SET_EMPTY and counts[2] == 0 and np.isnan(logits[:, 1, 0]).all()
    assert names[3] == "gridconv+" and counts[3] == 64 + 1
    assert names[5] == "mask" and counts[5] == 1
    sup_x = np.transpose(vol.sup, (0, 3, 1, 2))[None, :, None]
    for q in range(3):
        qq = np.transpose(qry[q], (2, 0, 1))[None]
        for l, (bg_ok, fg_mode) in enumerate([(True, "mask"), (False, "gridconv+"), (True, "mask"), (True, names[7])]):
            if bg_ok:
                ref, _, _, _ = O.alp_forward(qq, sup_x, (1.0 - fg[l])[None, :, None], "gridconv", 0.95, [2, 2], isval=True, val_wsize=2)
                np.testing.assert_allclose(logits[q, l, 0], ref[0, 0], atol=MAP_TOL, rtol=0)
            ref, _, _, _ = O.alp_forward(qq, sup_x, fg[l][None, :, None], fg_mode, 0.95, [2, 2], isval=True, val_wsize=2)
            np.testing.assert_allclose(logits[q, l, 1], ref[0, 0], atol=MAP_TOL, rtol=0)
    assert np.all(logits[1][~np.isnan(logits[1])] == 0.0)
    # prompts from those maps: NaN scores can never be foreground; everything still agrees with the oracle bit for bit
    hdr, recs = eng.run(_t(qry))
    with pytest.raises(RuntimeError):           # the reference raises for the empty 'gridconv' set; so does decode()
        eng.decode(hdr, recs)
    got = eng.decode(hdr, recs, on_empty_set="ignore")
    flat = logits.reshape(12, 2, h, w)
    for i in range(12):
        ref = O.coarse_to_prompts(flat[i][None], img, 256, use_cca=False, point_mode="both")
        s = got[i // 4][i % 4]
        assert s.empty == ref["empty"], i
        if not s.empty:
            assert np.array_equal(s.boxes, ref["bboxes"]) and np.array_equal(s.points, ref["points"])


def test_engine_rectangular_feature_maps_and_odd_channel_blocks():
    """h != w and C not a multiple of 64 (one partially filled k-block) through all three kernels"""
    h, w, C, img = 24, 40, 200, 256
    sup = synth.layer_norm(synth.gaussian_like(5, (1, h, w, C)))
    qry = (sup + 0.3 * synth.gaussian_like(6, (2, h, w, C))).astype(np.float32)
    fg = np.zeros((1, 1, h, w), np.float32)
    fg[0, 0, 6:18, 10:30] = 1.0
    eng = CoarseVolumeEngine((h, w), img, out_size=256, val_wsize=2, proto_grid_size=8)
    eng.set_support(_t(sup), _t(fg))
    logits = eng.match(_t(qry)).cpu().numpy()
    sup_x = np.transpose(sup, (0, 3, 1, 2))[None, :, None]
    for q in range(2):
        qq = np.transpose(qry[q], (2, 0, 1))[None]
        for ch, (mask, mode) in enumerate(((1.0 - fg[0], "gridconv"), (fg[0], "gridconv+"))):
            ref, _, _, _ = O.alp_forward(qq, sup_x, mask[None, :, None], mode, 0.95, [h // 8, w // 8], isval=True, val_wsize=2)
            np.testing.assert_allclose(logits[q, ch], ref[0, 0], atol=MAP_TOL, rtol=0)


def test_parted_volume_engine_vs_oracle_per_slice():
    """support-part protocol (dataloaders/common.py:228-252 + validation_protosam.py:352-388): every (slice, label)
    must equal the oracle run with the support slice of the part that slice falls in"""
    from protosam_b200.engine import PartedVolumeEngine, part_assign
    assert [part_assign(z, 10, 40, 3) for z in (5, 10, 19, 20, 29, 30, 40, 50)] == [0, 0, 0, 1, 1, 2, 2, 2]
    assert part_assign(7, 7, 7, 3) == 0
    h = w = 16
    C, img, L, npart, Q = 64, 128, 2, 3, 7
    sups = [synth.make_volume(60 + p, Q=1, L=L, C=C, h=h, w=w, img_size=img) for p in range(npart)]
    sup = np.concatenate([v.sup for v in sups], 0)                               # [npart,h,w,C]
    fg = np.stack([np.concatenate([v.fg[l] for v in sups], 0) for l in range(L)])  # [L,npart,h,w]
    qv = synth.make_volume(99, Q=Q, L=L, C=C, h=h, w=w, img_size=img)
    z_ids = [3, 8, 11, 14, 20, 27, 33]
    z_ranges = [(5, 29), (10, 33)]
    pe = PartedVolumeEngine(CoarseVolumeEngine((h, w), img, out_size=256, val_wsize=2, fg_mode="gridconv+"), npart)
    pe.set_support(_t(sup), _t(fg))
    hdr, recs, parts, logits = pe.run(_t(qv.qry), z_ids, z_ranges, return_logits=True)
    logits = logits.cpu().numpy()
    assert parts == [[part_assign(z, a, b, npart) for (a, b) in z_ranges] for z in z_ids]
    assert len({tuple(p) for p in parts}) >= 4                                    # several segments exercised
    got = pe.eng.decode(hdr, recs)
    for q in range(Q):
        qq = np.transpose(qv.qry[q], (2, 0, 1))[None]
        for l in range(L):
            p = parts[q][l]
            sx = np.transpose(sup[p:p + 1], (0, 3, 1, 2))[None, :, None]
            bg, _, _, _ = O.alp_forward(qq, sx, (1.0 - fg[l, p])[None, None, None], "gridconv", 0.95, [2, 2], isval=True, val_wsize=2)
            fgm, _, _, _ = O.alp_forward(qq, sx, fg[l, p][None, None, None], "gridconv+", 0.95, [2, 2], isval=True, val_wsize=2)
            np.testing.assert_allclose(logits[q * L + l], np.concatenate([bg, fgm], 1)[0], atol=MAP_TOL, rtol=0)
            ref = O.coarse_to_prompts(logits[q * L + l][None], img, 256, use_cca=False, point_mode="both")
            s = got[q][l]
            assert s.empty == ref["empty"]
            if not s.empty:
                assert np.array_equal(s.boxes, ref["bboxes"]) and np.array_equal(s.points, ref["points"])


@pytest.mark.parametrize("split", [False, True])
def test_graphed_volume_step_replays_equal_eager_run(split):
    """CUDA-graph replay of a volume (engine.GraphedVolumeStep) == the eager engine, bit for bit, also after the
    input buffers were refilled in place; split: the prompt stage replayed on the step's second stream"""
    from protosam_b200.engine import GraphedVolumeStep
    cfg = synth.CONFIGS["cfg2_chaos_mri"]
    vol = synth.make_volume(5, Q=3, L=2, C=128, h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    vol2 = synth.make_volume(6, Q=3, L=2, C=128, h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"])
    sup, fg, qry = _t(vol.sup), _t(vol.fg), _t(vol.qry)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gs = GraphedVolumeStep(eng, sup, fg, qry, split_streams=split)
        assert gs.n_kernels == 2 + 2 + 4            # kernel 1 (2 launches), prototype pack + fused GEMM, classify + blocks + components + compaction
        h1, r1 = gs.launch()
        gs.join()
        h1, r1 = h1.clone(), r1.clone()
        sup.copy_(_t(vol2.sup)); fg.copy_(_t(vol2.fg)); qry.copy_(_t(vol2.qry))
        h2, r2 = gs.launch()
        gs.join()
        h2, r2 = h2.clone(), r2.clone()
    s.synchronize()
    for v, (h, r) in ((vol, (h1, r1)), (vol2, (h2, r2))):
        e2 = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"])
        e2.set_support(_t(v.sup), _t(v.fg))
        he, re_ = e2.run(_t(v.qry))
        assert torch.equal(h, he) and torch.equal(r, re_)
    # world == 1: run_sharded degenerates to run, also with an asynchronous handle
    e2.set_support(_t(vol.sup), _t(vol.fg))
    ha, ra = e2.run_sharded(_t(vol.qry), 3, async_op=True).result()
    assert torch.equal(ha, h1) and torch.equal(ra, r1)


@pytest.mark.parametrize("fg_mode", ["auto_fg", "mask"])
def test_engine_multi_shot_matches_reference_rule(fg_mode):
    """n_shots > 1 (grid_proto_fewshot.py:239-270): background prototypes from all shots at once, foreground once per
    shot (that shot alone: its windows, its mode decision, its single global row), element-wise max over the shots"""
    h = w = 16
    C, img, L, S, Q = 64, 128, 2, 3, 2
    sup = synth.layer_norm(synth.gaussian_like(41, (S, h, w, C)))
    qry = (np.roll(sup[:Q], 1, 1) + 0.3 * synth.gaussian_like(42, (Q, h, w, C))).astype(np.float32)
    fg = np.zeros((L, S, h, w), np.float32)
    fg[0, 0, 2:8, 2:8] = 1; fg[0, 1, 6:12, 5:11] = 1; fg[0, 2, 9, 9] = 1      # shot 2 of label 0: single pixel -> 'mask'
    fg[1, 0, 10:14, 1:5] = 1; fg[1, 2, 3:9, 8:14] = 1                          # shot 1 of label 1: empty mask
    eng = CoarseVolumeEngine((h, w), img, out_size=256, val_wsize=2, proto_grid_size=8, fg_mode=fg_mode)
    pr = eng.set_support(_t(sup), _t(fg))
    assert pr["protos"].shape[0] == L * (1 + S)
    logits = eng.match(_t(qry)).cpu().numpy().reshape(Q, L, 2, h, w)
    eff = [_lib.MODE_NAMES[int(e)] for e in pr["eff_modes"].cpu().numpy()]
    ks = [h // 8, w // 8]
    for l in range(L):
        bg_mask = (1.0 - fg[l])[None, :, None]                                    # [1,S,1,h,w], all shots
        sx_all = np.transpose(sup, (0, 3, 1, 2))[None, :, None]
        for q in range(Q):
            qq = np.transpose(qry[q], (2, 0, 1))[None]
            ref_bg, _, _, _ = O.alp_forward(qq, sx_all, bg_mask, "gridconv", 0.95, ks, isval=True, val_wsize=2)
            np.testing.assert_allclose(logits[q, l, 0], ref_bg[0, 0], atol=MAP_TOL, rtol=0)
            per_shot = []
            for s in range(S):
                if fg_mode == "mask":
                    mode = "mask"
                else:   # the caller's rule on that shot's mask (grid_proto_fewshot.py:250-256)
                    pooled = O.get_prototypes(np.transpose(sup[s:s + 1], (0, 3, 1, 2)), fg[l, s][None, None], "gridconv+", ks, 0.95)["pooled"]
                    mode = "gridconv+" if (pooled >= 0.95).any() else "mask"
                assert eff[l * (1 + S) + 1 + s] == mode
                sx = np.transpose(sup[s:s + 1], (0, 3, 1, 2))[None, :, None]
                r, _, _, _ = O.alp_forward(qq, sx, fg[l, s][None, None, None], mode, 0.95, ks, isval=True, val_wsize=2)
                per_shot.append(r[0, 0])
            np.testing.assert_allclose(logits[q, l, 1], np.max(np.stack(per_shot), 0), atol=MAP_TOL, rtol=0)
    got = eng.decode(*eng.run(_t(qry)))
    assert len(got) == Q and len(got[0]) == L


@pytest.mark.parametrize("point_mode", ["conf", "centroid", "both"])
@pytest.mark.parametrize("orig", [(1024, 1024), (672, 512)])
def test_records_to_sam_tensors_match_reference_transforms(point_mode, orig):
    """device prompt tensors == ResizeLongestSide.apply_coords / apply_boxes (segment_anything/utils/transforms.py:40-62)
    followed by torch.as_tensor(dtype=float) (predictor.py:143-150), applied to the reference-format prompts"""
    low = (synth.gaussian_like(77, (3, 2, 24, 24)) * 6).astype(np.float32)
    low[2] = -5.0 * np.abs(low[2]) * np.array([-1.0, 1.0], np.float32)[:, None, None]     # image 2: no foreground
    hdr, recs = ops.coarse_to_prompts(_t(low), 96, 256, use_cca=False, max_cc=64)
    pts, labs, boxes = ops.records_to_sam(hdr, recs, point_mode, original_size=orig, target_length=1024)
    H, R = ops.decode_headers(hdr), ops.decode_records(recs)
    old_h, old_w = orig
    scale = 1024 * 1.0 / max(old_h, old_w)
    new_h, new_w = int(old_h * scale + 0.5), int(old_w * scale + 0.5)
    pts, labs, boxes = pts.cpu().numpy(), labs.cpu().numpy(), boxes.cpu().numpy()
    assert int(H["n_rec"][2]) == 0 and not pts[2].any() and not boxes[2].any()
    for i in range(3):
        sp = PR.prompts_from_records(H[i], R[i], False, point_mode)
        n = int(H["n_rec"][i])
        if n == 0:
            continue
        coords = sp.points.astype(float).copy()
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        assert np.array_equal(pts[i, :n], coords.astype(np.float32))
        b = sp.boxes.reshape(-1, 2, 2).astype(float)
        b[..., 0] = b[..., 0] * (new_w / old_w)
        b[..., 1] = b[..., 1] * (new_h / old_h)
        assert np.array_equal(boxes[i, :n], b.reshape(-1, 4).astype(np.float32))
        assert (labs[i, :n] == 1).all() and not labs[i, n:].any()


def test_tokens_to_features_matches_aten_bilinear():
    """the DINOv2 token hand-off of BASELINE config 1 (18x18 tokens -> 32x32 features, grid_proto_fewshot.py:96-98)"""
    tok = synth.gaussian_like(3, (2, 18 * 18, 48))
    got = ops.tokens_to_features(_t(tok)).cpu().numpy()                        # [2,32,32,48] channels-last
    ref = O.upsample_bilinear(np.transpose(tok.reshape(2, 18, 18, 48), (0, 3, 1, 2)), 32)   # ATen-CPU restatement
    assert got.shape == (2, 32, 32, 48) and np.array_equal(np.transpose(got, (0, 3, 1, 2)), ref)
    big = _t(synth.gaussian_like(4, (1, 37 * 37, 16)))
    assert ops.tokens_to_features(big).data_ptr() == big.data_ptr()              # >= 32x32 tokens: a view, no copy


# ------------------------------------------------------------------------------ compact records

def test_compact_records_equal_dense_records():
    """psam_compact_records: headers + live records only (what gathers and host copies move) decode to the same prompts as
    the dense [n, max_cc] arrays; too small a capacity is flagged, never silently truncated."""
    cfg = synth.CONFIGS["cfg2_chaos_mri"]
    vol = synth.make_volume(1234, Q=3, L=2, C=cfg["C"], h=cfg["h"], w=cfg["w"], img_size=cfg["img_size"])
    eng = CoarseVolumeEngine((cfg["h"], cfg["w"]), cfg["img_size"], val_wsize=cfg["ws"])
    eng.set_support(_t(vol.sup), _t(vol.fg))
    logits = eng.match(_t(vol.qry))
    n = logits.shape[0]
    hdr, recs, packed = eng.prompts_from_logits(logits, n_alloc=n + 5, return_packed=True)
    cap = (n + 5) * eng.recs_per_image
    assert packed.numel() == ops.packed_bytes(n + 5, cap) == _lib.load().psam_packed_bytes(n + 5, cap)
    H, R = ops.decode_packed(packed, n + 5, cap)
    Hd, Rd = ops.decode_headers(hdr), ops.decode_records(recs)
    assert len(H) == n and len(R) == int(Hd["n_rec"].sum()) > 0
    for i in range(n):
        k, a = int(Hd["n_rec"][i]), int(H["reserved"][i])
        assert all(np.array_equal(H[f][i], Hd[f][i]) for f in Hd.dtype.names if f != "reserved")
        assert R[a: a + k].tobytes() == Rd[i, :k].tobytes()
    ph, _, pr = ops.split_packed(packed, n + 5, cap)
    a = eng.decode(hdr, recs)
    b = eng.decode(ph[:n], pr[: len(R)])
    for sa, sb in zip(sum(a, []), sum(b, [])):
        assert sa.empty == sb.empty and (sa.empty or (np.array_equal(sa.boxes, sb.boxes) and np.array_equal(sa.points, sb.points)))
    small = ops.compact_records(hdr, recs, n_alloc=n, capacity=1)
    with pytest.raises(RuntimeError):
        ops.decode_packed(small, n, 1)


def test_engine_reports_more_components_than_max_cc():
    """more components than max_cc: the reference would prompt SAM for all of them, so decode() raises instead of
    returning a truncated list (IMG_CC_TRUNCATED)"""
    low = (synth.gaussian_like(5, (1, 2, 48, 48)) * 6).astype(np.float32)          # speckle: hundreds of components
    eng = CoarseVolumeEngine((48, 48), 672, max_cc=8)
    eng.n_labels = 1
    hdr, recs = eng.prompts_from_logits(_t(low))
    H = ops.decode_headers(hdr)[0]
    assert int(H["ncc"]) > 8 and int(H["flags"]) & _lib.IMG_CC_TRUNCATED
    with pytest.raises(RuntimeError):
        eng.decode(hdr, recs)
