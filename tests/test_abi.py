"""The C-ABI boundary: libpsam_b200.so loads without a GPU and exports every symbol
include/psam_b200.h declares; the ctypes table in protosam_b200/_lib.py matches the header."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "psam_b200.h")


@pytest.fixture(scope="module")
def lib():
    from protosam_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):       # the driver normally runs build() first
        subprocess.run(["make", "-C", os.path.join(ROOT, "protosam_b200", "csrc"), "-j4"], check=True)
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psam_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/psam_b200.h but not exported"


def test_ctypes_table_covers_header(lib):
    from protosam_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes table {len(args)}"


def test_abi_version_and_struct_layout(lib):
    from protosam_b200 import ops
    assert lib.psam_abi_version() == 3
    assert ops.REC_DTYPE.itemsize == 96 and ops.HDR_DTYPE.itemsize == 64
    assert ops.REC_DTYPE.fields["centroid"][1] == 48 and ops.REC_DTYPE.fields["conf"][1] == 64
    assert ops.HDR_DTYPE.fields["bg_centroid"][1] == 40


def test_workspace_queries_need_no_gpu(lib):
    assert lib.psam_alp_prototypes_workspace(8, 1, 768, 37, 37, 2, 2) > 8 * 37 * 768 * 4
    assert lib.psam_alp_prototypes_workspace(0, 1, 768, 37, 37, 2, 2) == 0
    assert lib.psam_coarse_to_prompts_workspace(4, 1024, 65536, 64) > 4 * 1024 * 1024 * 4


def test_argument_errors_are_reported_not_crashed(lib):
    rc = lib.psam_alp_match(None, 0, 0, 1, 1, 4, None, 1, None, None, 1, None, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.psam_last_error()


def test_new_entry_points_check_their_arguments(lib):
    """psam_topk_points, psam_peer_*, psam_match_reserve_sms: argument errors come back as codes + messages (no GPU needed:
    the checks run before anything is enqueued)"""
    one = ctypes.c_void_p(16)                      # a non-null, 16-byte aligned dummy pointer: never dereferenced
    rc = lib.psam_topk_points(one, one, 0, one, one, 1, 64, 4, 0, 65, one, one, one, 1 << 30, None)
    assert rc == -1 and b"k = 65" in lib.psam_last_error()
    rc = lib.psam_topk_points(one, one, 0, one, one, 1, 64, 4, 0, 3, one, one, None, 0, None)
    assert rc == -2 and b"workspace" in lib.psam_last_error()
    assert lib.psam_topk_points_workspace(2, 8, 3) >= 2 * 8 * 64 * 3 * 8
    assert lib.psam_peer_region_bytes(1000) == 4096 + 1024
    rc = lib.psam_peer_put(one, 32, 32, one, 1, 0, 0, one, None)             # world < 2
    assert rc == -1 and b"world" in lib.psam_last_error()
    rc = lib.psam_peer_put(one, 24, 32, one, 2, 0, 1, one, None)             # not a multiple of 16 bytes
    assert rc == -1
    rc = lib.psam_peer_push_table(one, 8, 10, 6, 0, 0, one, 2, 0, one, None)   # C % 4, integer arrays inside the rows
    assert rc == -1 and b"geometry" in lib.psam_last_error()
    rc = lib.psam_peer_recv_table(one, 8, 10, 8, 8 * 10 * 8 * 4, 96, one, 2, 1, 1, one, None)   # src == rank
    assert rc == -1
    prev = lib.psam_match_reserve_sms(5)
    assert lib.psam_match_reserve_sms(-1) == 5
    lib.psam_match_reserve_sms(prev)
    assert lib.psam_match_reserve_sms(-1) == prev


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under protosam_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "protosam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libpsam_oracle" not in txt and "oracle._build" not in txt, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from protosam_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from protosam_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.upsample_softmax(torch.zeros(1, 2, 8, 8), 16, 32)


def test_medsam_box_handoff_matches_reference_formula():
    """models/ProtoMedSAM.py:197-200"""
    from protosam_b200 import prompts
    box = np.array([[10, 20, 300, 400], [0, 0, 671, 671]], np.int64)
    out = prompts.medsam_boxes(box, 672, 672)
    assert out.dtype == np.float64 and np.array_equal(out, box / np.array([672, 672, 672, 672]) * 1024)
