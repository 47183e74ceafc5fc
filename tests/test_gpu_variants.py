"""The optional prompt variants (negative points, mask prompts, coarse_pred_only confidence; SURVEY.md section 8(f)
rank 3) and the ProtoMedSAM box path (section 8(a) a15) on the GPU, against fixtures the unmodified reference
produced (tests/golden/variants.npz, oracle/make_golden.py::gen_variants) and against the oracle."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from protosam_b200 import ops, prompts as PR, synth  # noqa: E402
from protosam_b200.engine import CoarseVolumeEngine  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _g():
    return np.load(os.path.join(GOLD, "variants.npz"))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


NAMES = list(_g()["names"])


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("dev", [0, 1])
@pytest.mark.parametrize("use_cca", [False, True])
def test_negative_points_vs_reference_golden(name, dev, use_cca):
    """the points / labels every SamPredictor.predict call receives with use_neg_points=True, for both `.cpu()`
    semantics of the reference (dev = 1: its CUDA path)"""
    g = _g()
    key = f"{name}/neg_dev{dev}_cca{int(use_cca)}"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        sp = PR.coarse_to_prompts(_t(g[f"{name}/low"]), int(g[f"{name}/S"]), 1024, use_cca=use_cca, point_mode="both",
                                  max_cc=1024, use_neg_points=True, host_aliasing=(dev == 0))[0]
    calls = sp.predict_calls()
    assert len(calls) == int(g[f"{key}/ncalls"])
    for i, c in enumerate(calls):
        want_lab = g[f"{key}/point_labels"][i]
        n = int((want_lab >= 0).sum())
        assert c["point_coords"].shape == (n, 2)
        assert np.array_equal(c["point_coords"], g[f"{key}/points"][i][:n])
        assert np.array_equal(c["point_labels"], want_lab[:n])


@pytest.mark.parametrize("name", NAMES)
def test_mask_prompts_vs_reference_golden(name):
    g = _g()
    sp = PR.coarse_to_prompts(_t(g[f"{name}/low"]), int(g[f"{name}/S"]), 1024, use_cca=False, point_mode="both",
                              max_cc=1024, use_mask=True)[0]
    calls = sp.mask_predict_calls()
    assert len(calls) == int(g[f"{name}/mask/ncalls"])
    if calls:
        m = np.stack([c["mask_input"] for c in calls])
        assert m.shape == (len(calls), 1, 256, 256) and str(m.dtype) == str(g[f"{name}/mask/dtype"])
        assert np.array_equal(np.unique(m), g[f"{name}/mask/values"])
        assert np.array_equal(np.packbits(m == 10), g[f"{name}/mask/fg_bits"])


@pytest.mark.parametrize("name", NAMES)
def test_coarse_pred_only_confidence_vs_reference_golden(name):
    """ProtoSAM(coarse_pred_only=True): logits at the ALPNet image size (256 / 518: any `out` is accepted), the
    confidence of util/utils.py:429-434, and cca(..., return_conf=True) with use_cca"""
    g = _g()
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    p_fg, bits, _ = ops.upsample_softmax(_t(low), S, S)                 # FewShotSeg's single upsample to S x S
    wpr = (S + 31) // 32
    mask = np.unpackbits(bits[0].cpu().numpy().view(np.uint8), bitorder="little").reshape(S, wpr * 32)[:, :S]
    assert np.array_equal(np.packbits(mask), g[f"{name}/coarse_cca0/pred_bits"])
    conf = float(ops.confidence(p_fg)[0].item())
    assert conf == pytest.approx(float(g[f"{name}/coarse_cca0/conf"]), rel=2e-6, abs=1e-7)
    logits_S = _t(O.upsample_bilinear(low, S))
    assert PR.get_confidence_from_logits(logits_S) == pytest.approx(float(g[f"{name}/coarse_cca0/conf"]), rel=2e-6, abs=1e-7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        kept, c = PR.cca(mask.astype(np.uint8), logits_S, return_conf=True)
    assert np.array_equal(np.packbits(kept), g[f"{name}/coarse_cca1/pred_bits"])
    assert float(c) == pytest.approx(float(g[f"{name}/coarse_cca1/conf"]), rel=1e-5, abs=1e-9)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("use_cca", [False, True])
def test_protomedsam_boxes_vs_reference_golden(name, use_cca):
    """engine variant 'medsam': boxes handed to medsam_inference; confidences of a softmax applied twice"""
    g = _g()
    low, S = g[f"{name}/low"], int(g[f"{name}/S"])
    h = low.shape[-1]
    eng = CoarseVolumeEngine((h, h), S, out_size=1024, use_cca=use_cca, max_cc=1024, variant="medsam")
    eng.n_labels = 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        sp = eng.decode(*eng.prompts_from_logits(_t(low)))[0][0]
    key = f"{name}/medsam_cca{int(use_cca)}"
    if sp.empty:
        assert int(g[f"{key}/ncalls"]) == 0
        return
    boxes = PR.medsam_boxes(sp.boxes, 1024, 1024)
    assert boxes.dtype == g[f"{key}/boxes"].dtype and np.array_equal(boxes, g[f"{key}/boxes"])
    if not use_cca:
        want = g[f"{name}/medsam_conf"]
        got = np.array([sp.conf[k] for k in sorted(sp.conf)], dtype=np.float64)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    ref = O.coarse_to_prompts_medsam(low, S, 1024, use_cca=use_cca)
    assert np.array_equal(boxes, ref["boxes_1024"])


def test_medsam_every_pixel_variant_and_function_level_cca():
    """prob_mode 'softmax_twice' of the every-pixel kernel equals the oracle's double softmax bit for bit, the engine
    variant writes the same values where it writes, and the function-level cca() drop-in handed PROBABILITIES (what
    ProtoMedSAM passes, models/ProtoMedSAM.py:178-185) picks the oracle's component."""
    low = (synth.gaussian_like(77, (2, 2, 37, 37)) * 7).astype(np.float32)
    p_full, bits_full, probs2 = ops.upsample_softmax(_t(low), 518, 1024, want_probs2=True, prob_mode="softmax_twice")
    p_eng, bits, _, wstat = ops.upsample_softmax(_t(low), 518, 1024, fg_only=True, want_wstat=True, prob_mode="softmax_twice")
    assert torch.equal(bits, bits_full)
    for i in range(2):
        _, p, pred = O.coarse_logits_to_probs(low[i:i + 1], 518, 1024)
        assert np.array_equal(probs2[i].cpu().numpy(), p[0])
        assert np.array_equal(p_full[i].cpu().numpy(), O.softmax2(p)[0, 1])
    written = p_eng != 0
    assert written.any() and torch.equal(p_eng[written], p_full[written])
    for use_cca in (False, True):
        a = ops.components(bits, p_eng, use_cca=use_cca, max_cc=4096, wstat=wstat)
        b = ops.components(bits_full, p_full, use_cca=use_cca, max_cc=4096)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    _, p, pred = O.coarse_logits_to_probs(low[:1], 518, 1024)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        cc = PR.cca(pred, _t(p), return_cc=True)
    ref = O.cca(pred, O.softmax2(p)[0, 1], return_cc=True)
    assert cc[0] == ref[0] and np.array_equal(cc[1], ref[1]) and np.array_equal(cc[2], ref[2])


def test_function_level_negative_points_and_mask_inputs_match_oracle():
    low = (synth.gaussian_like(31, (1, 2, 24, 24)) * 6).astype(np.float32)
    lg, p, pred = O.coarse_logits_to_probs(low, 256, 512)
    cc, _ = PR.get_connected_components(pred, _t(lg), return_conf=True)
    rcc, _ = O.get_connected_components(pred, p[0, 1], return_conf=True)
    for alias in (False, True):
        pts, labels, neg, neg_labels = PR.get_sam_input_points(cc, None, get_neg_points=True, l=1, point_mode="conf",
                                                               host_aliasing=alias)
        ref = O.get_neg_points(rcc, p, host_aliasing=alias)
        assert len(neg) == len(ref) == rcc[0] - 1 and len(neg_labels) == len(neg)
        for a, b in zip(neg, ref):
            assert (a is None and b is None) or np.array_equal(a, b)
    m = PR.sam_mask_inputs(cc)
    rm, ids = O.sam_mask_inputs(rcc)
    assert np.array_equal(m, rm)
    fm, fids = PR.get_sam_input_mask(cc)
    assert fm.shape == (rcc[0] - 1, 512, 512) and np.array_equal(fids, ids)
