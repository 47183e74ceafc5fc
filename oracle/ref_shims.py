"""Import the UNMODIFIED reference (levayz/ProtoSAM) on a CPU-only box.

TEST INFRASTRUCTURE ONLY.  Nothing under ``protosam_b200/`` may import this
file.  It is used by ``oracle/make_golden.py`` (fixture generation, in the
build container where ``/root/reference`` is mounted) and by CPU tests that
are skipped when the reference tree is absent (it never travels to the GPU
box).

Three shims, none of which edits a reference file:

1. ``matplotlib`` / ``matplotlib.pyplot`` / ``kneed`` are replaced by empty
   stub modules (imported at models/alpmodule.py:11, models/ProtoSAM.py:5,
   util/utils.py:10-11; only called under ``debug``).
2. ``torch.Tensor.cuda`` becomes the identity while a reference call runs, so
   the hard-coded ``.cuda()`` in ``safe_norm`` (models/alpmodule.py:16,184)
   works on CPU.
3. ``torch.hub.load`` returns a stub DINOv2 exposing ``forward_features``
   (models/grid_proto_fewshot.py:55-72,90-91) and ``ProtoSAM.get_sam`` is
   replaced by a capturing predictor (models/ProtoSAM.py:205-220) so no
   checkpoint is needed.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PSAM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "alpmodule.py"))


def _install_stub_modules() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "kneed"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


@contextlib.contextmanager
def cpu_cuda_identity():
    """Make ``Tensor.cuda()`` a no-op for the duration of a reference call."""
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


_ALP = None


def load_alpmodule():
    """The reference models/alpmodule.py as an isolated module object."""
    global _ALP
    if _ALP is None:
        _install_stub_modules()
        spec = importlib.util.spec_from_file_location(
            "ref_alpmodule", os.path.join(REFERENCE_ROOT, "models", "alpmodule.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ALP = mod
    return _ALP


class StubDino(nn.Module):
    """[B,3,14k,14k] -> {"x_norm_patchtokens": [B,k*k,C]} (LayerNorm'd like DINOv2)."""

    def __init__(self, C: int):
        super().__init__()
        self.C = C
        self.proj = nn.Conv2d(3, C, 14, 14)

    def forward_features(self, x):
        t = self.proj(x).flatten(2).transpose(1, 2)
        return {"x_norm_patchtokens": nn.functional.layer_norm(t, (self.C,))}


_DINO_DIMS = {"dinov2_vitb14": 768, "dinov2_vitl14": 1024, "dinov2_vitl14_reg": 1024}


class CapturingPredictor:
    """Stands in for SamPredictor: records every predict() call's prompts."""

    def __init__(self):
        self.calls = []
        self.image = None

    def set_image(self, image):
        self.image = image

    def predict(self, **kw):
        self.calls.append({k: (None if v is None else np.array(v)) if k in
                           ("point_coords", "point_labels", "box", "mask_input") else v
                           for k, v in kw.items()})
        h, w = self.image.shape[:2]
        return (np.zeros((1, h, w), dtype=bool), np.ones((1,), dtype=np.float32), None)


_PIPE = None


def load_pipeline():
    """(FewShotSeg, ProtoSAM module, util.utils) of the reference."""
    global _PIPE
    if _PIPE is None:
        _install_stub_modules()
        for p in (os.path.join(REFERENCE_ROOT, "models"), REFERENCE_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        torch.hub.load = lambda repo, name, **k: StubDino(_DINO_DIMS[name])
        gpf = importlib.import_module("models.grid_proto_fewshot")
        PS = importlib.import_module("models.ProtoSAM")
        uu = importlib.import_module("util.utils")
        from models.segment_anything.utils.transforms import ResizeLongestSide

        def get_sam(self, checkpoint_path, use_sam_trans):
            self.predictor = CapturingPredictor()
            if use_sam_trans:
                t = ResizeLongestSide(1024)
                t.pixel_mean = torch.tensor([0, 0, 0]).view(3, 1, 1)
                t.pixel_std = torch.tensor([1, 1, 1]).view(3, 1, 1)
                self.sam_trans = t
            else:
                self.sam_trans = None

        PS.ProtoSAM.get_sam = get_sam
        _PIPE = (gpf, PS, uu)
    return _PIPE


_MED = None


def load_protomedsam():
    """models/ProtoMedSAM.py with the MedSAM network replaced by a stub: ``medsam_inference`` records the boxes it is
    handed (the end of the path this repo owns) and returns an empty mask, so no checkpoint is needed."""
    global _MED
    if _MED is None:
        load_pipeline()
        PM = importlib.import_module("models.ProtoMedSAM")

        class _Enc(nn.Module):
            def forward(self, x):
                return torch.zeros(1, 256, 64, 64)

        class _Med(nn.Module):
            def __init__(self):
                super().__init__()
                self.image_encoder = _Enc()

        def get_sam(self, checkpoint_path):
            self.medsam = _Med()
            self.captured_boxes = []

        def medsam_inference(self, img_embed, box_1024, H, W, query_label=None):
            self.captured_boxes.append(np.array(box_1024))
            return np.zeros((H, W), np.uint8), 1.0

        PM.ProtoMedSAM.get_sam = get_sam
        PM.ProtoMedSAM.medsam_inference = medsam_inference
        _MED = PM
    return _MED


@contextlib.contextmanager
def device_copy_semantics():
    """On the reference's CUDA path ``tensor.cpu()`` COPIES; on a CPU tensor it returns the tensor itself, so in-place
    edits of the result alias the source (models/ProtoSAM.py:363-364 vs :414).  Inside this context ``.cpu()`` clones,
    which reproduces what the CUDA path computes."""
    orig = torch.Tensor.cpu
    torch.Tensor.cpu = lambda self, *a, **k: self.clone()
    try:
        yield
    finally:
        torch.Tensor.cpu = orig


class FixedLogitsCoarseModel:
    """Coarse-model stub for ProtoSAM.forward: returns pre-computed logits."""

    def __init__(self, logits):
        self.logits = logits

    def __call__(self, coarse_model_input):
        return self.logits


class _NullInput:
    def set_query_images(self, q):
        pass
