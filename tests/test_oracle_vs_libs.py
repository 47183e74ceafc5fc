"""Each third-party primitive the reference path leans on (ATen CPU, numpy, OpenCV),
restated in oracle/psam_oracle.c, checked bit-for-bit against the library itself."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from protosam_b200 import synth

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("ih,oh", [(37, 518), (518, 1024), (32, 256), (256, 1024), (48, 672), (672, 1024),
                                   (73, 1024), (24, 24)])
def test_bilinear_bit_exact(ih, oh):
    x = torch.from_numpy(synth.gaussian_like(ih * 1000 + oh, (1, 2, ih, ih)) * 8)
    ref = F.interpolate(x, size=(oh, oh), mode="bilinear").numpy()
    assert np.array_equal(O.upsample_bilinear(x.numpy(), oh), ref)


def test_bilinear_rectangular():
    x = torch.from_numpy(synth.gaussian_like(5, (1, 2, 21, 29)) * 8)
    ref = F.interpolate(x, size=(300, 170), mode="bilinear").numpy()
    assert np.array_equal(O.upsample_bilinear(x.numpy(), (300, 170)), ref)


@pytest.mark.parametrize("scale", [0.01, 1.0, 8.0, 30.0, 120.0])
def test_softmax2_bit_exact(scale):
    l = torch.from_numpy(synth.gaussian_like(int(scale * 100) + 3, (1, 2, 1024, 1024)) * scale)
    assert np.array_equal(O.softmax2(l.numpy()), l.softmax(1).numpy())


@pytest.mark.parametrize("ws", [2, 3, 4, 5, 7, 8])
def test_avg_pool_mask_bit_exact(ws):
    y = synth.uniform(ws, (2, 1, 37, 37))
    L = O.lib()
    out = np.zeros((2, 37 // ws, 37 // ws), np.float32)
    yc = np.ascontiguousarray(y.reshape(2, 37, 37))
    L.psamo_pool_mask(yc.ctypes.data_as(ctypes.c_void_p), 2, 37, 37, ws, ws, out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out, F.avg_pool2d(torch.from_numpy(y), ws).numpy()[:, 0])


@pytest.mark.parametrize("n", [1 << 20, 518 * 518, 1000, 129, 128, 77, 8, 5])
def test_pairwise_sum_matches_numpy(n):
    L = O.lib()
    for t in range(3):
        a = synth.uniform(n + t, (n,))
        a = a * (synth.uniform(n + t + 99, (n,)) < 0.3)
        got = L.psamo_pairwise_sum_f32(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n))
        assert np.float32(got) == a.sum()


def test_cc_sums_match_numpy_per_component():
    rng = np.random.default_rng(0)
    low = rng.random((40, 40)) < 0.45
    mask = np.kron(low, np.ones((8, 8), bool))[:300, :317]
    p = synth.uniform(3, mask.shape)
    n, labels, _, _ = O.connected_components(mask)
    sums = O.cc_sums(p, labels, n)
    for j in range(1, n):
        assert sums[j] == (p.flatten() * (labels == j).flatten()).sum()


@pytest.mark.parametrize("H,W,density,seed", [(64, 64, 0.5, 0), (104, 131, 0.4, 1), (257, 300, 0.55, 2),
                                              (1024, 1024, 0.45, 3), (1024, 1024, 0.6, 4), (33, 1, 0.5, 5),
                                              (1, 40, 0.5, 6), (1023, 1021, 0.5, 7)])
def test_ccl_matches_opencv(H, W, density, seed):
    """labels (numbering included), stats and centroids equal cv2's, pixel noise masks."""
    rng = np.random.default_rng(seed)
    mask = (rng.random((H, W)) < density).astype(np.uint8)
    n, lab, st, ce = cv2.connectedComponentsWithStats(mask, connectivity=8)
    n2, lab2, st2, ce2 = O.connected_components(mask)
    assert n == n2 and np.array_equal(lab, lab2)
    assert np.array_equal(st, st2) and np.array_equal(ce, ce2)


def test_ccl_blobs_matches_opencv():
    rng = np.random.default_rng(11)
    for t in range(6):
        low = (rng.random((24, 24)) < 0.3).astype(np.float32)
        up = F.interpolate(torch.from_numpy(low)[None, None], size=(1024, 1024), mode="bilinear")[0, 0].numpy()
        mask = (up > 0.4).astype(np.uint8)
        n, lab, st, ce = cv2.connectedComponentsWithStats(mask, connectivity=8)
        n2, lab2, st2, ce2 = O.connected_components(mask)
        assert n == n2 and np.array_equal(lab, lab2) and np.array_equal(st, st2) and np.array_equal(ce, ce2)


def test_topk1_tie_behaviour_matches_torch():
    """torch.topk(v, 1) picks different equal maxima for n < 64 (nth_element) and n >= 64
    (partial_sort); the oracle replays both."""
    L = O.lib()
    L.psamo_topk1_pos.restype = ctypes.c_int
    rng = np.random.default_rng(5)
    for n in list(range(1, 70)) + [100, 1000]:
        for t in range(40):
            levels = rng.integers(1, 5)
            v = (rng.integers(0, levels + 1, n) / levels).astype(np.float32)
            if t % 3 == 0:
                v = rng.random(n).astype(np.float32)
            ref = int(torch.topk(torch.from_numpy(v), 1).indices[0])
            got = L.psamo_topk1_pos(v.ctypes.data_as(ctypes.c_void_p), n)
            assert got == ref, (n, t, v.tolist())


def test_topk_any_k_tie_behaviour_matches_torch():
    """torch.topk(v, k): partial_sort when k * 64 <= n, nth_element + sort otherwise; with many equal values the order
    among ties is a function of libstdc++'s algorithms, which the oracle replays (get_most_conf_points, k > 1)."""
    L = O.lib()
    L.psamo_topk_pos.restype = ctypes.c_int
    rng = np.random.default_rng(11)
    cases = [(n, k) for n in (1, 2, 3, 4, 5, 7, 16, 17, 18, 33, 63, 64, 65, 127, 128, 129, 200, 319, 320, 321, 640, 1000, 5000)
             for k in (1, 2, 3, 5, 10, 16, 17, 40) if k <= n]
    for n, k in cases:
        for t in range(12):
            levels = int(rng.integers(1, 6))
            v = (rng.integers(0, levels + 1, n) / levels).astype(np.float32)
            if t % 4 == 0:
                v = rng.random(n).astype(np.float32)
            if t % 4 == 1:                                   # saturated probabilities: mostly exactly 1.0
                v = np.where(rng.random(n) < 0.8, 1.0, rng.random(n)).astype(np.float32)
            ref = torch.topk(torch.from_numpy(v), k).indices.numpy()
            got = np.zeros(k, np.int32)
            assert L.psamo_topk_pos(v.ctypes.data_as(ctypes.c_void_p), n, k, got.ctypes.data_as(ctypes.c_void_p)) == 0
            assert np.array_equal(got, ref), (n, k, t)
    got = np.zeros(4, np.int32)
    v = np.ones(3, np.float32)
    assert L.psamo_topk_pos(v.ctypes.data_as(ctypes.c_void_p), 3, 4, got.ctypes.data_as(ctypes.c_void_p)) == -1


def test_nearest_resize_matches_torch():
    m = synth.ellipse_mask(5, 518)
    for hw in (37, 32, 48, 73):
        ref = F.interpolate(torch.from_numpy(m)[None, None], size=(hw, hw), mode="nearest")[0, 0].numpy()
        assert np.array_equal(ref, synth.nearest_resize(m, hw, hw))
