"""Drop-in replacement of the reference ALP module (models/alpmodule.py:21-198).

Same constructor, attributes, ``forward`` signature and output shapes as the reference's
``MultiProtoAsConv``, so ``FewShotSeg.get_cls`` (models/grid_proto_fewshot.py:117) can
construct it and ``FewShotSeg.forward`` (:239-240, 258-259) can call it unchanged.  The
arithmetic runs in libpsam_b200.so (kernels 1 and 2); nothing is computed in PyTorch and there
is no CPU fallback.

Differences that are deliberate:
  * inference only: raises if autograd would have to flow through (training and ``alignLoss``
    stay on the reference module);
  * ``proto_grid`` (viz-only 4th output) is a CUDA tensor; the reference builds it on the CPU;
  * an invalid ``mode`` raises ``ValueError`` up front (the reference trips an
    ``UnboundLocalError`` in get_prototypes before reaching its own ValueError at :93-94);
  * ``forward`` never synchronises the host: the reference's error for an empty 'gridconv' set is raised by
    ``raise_if_empty()`` (or inside forward when ``check_empty = True``), see ``__init__``.  ``vis_sim=True`` (debug
    visualisation) does read the prototype count back to slice ``raw_local_sims``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib, ops

VALID_MODES = ("mask", "gridconv", "gridconv+")


class MultiProtoAsConv(nn.Module):
    def __init__(self, proto_grid, feature_hw, embed_dim=768, use_attention=False, upsample_mode="bilinear"):
        """Mirrors models/alpmodule.py:22-55: no parameters, no buffers (a state_dict saved from the
        reference loads with strict=True, models/grid_proto_fewshot.py:41-42)."""
        super().__init__()
        self.feature_hw = feature_hw
        self.proto_grid = proto_grid
        self.upsample_mode = upsample_mode
        kernel_size = [ft_l // grid_l for ft_l, grid_l in zip(feature_hw, proto_grid)]
        self.kernel_size = kernel_size
        print(f"MultiProtoAsConv: kernel_size: {kernel_size}")
        self.avg_pool_op = nn.AvgPool2d(kernel_size)
        if use_attention:
            # dead weights in the reference too (never touched by forward); kept so that
            # checkpoints trained with use_attention=True still load strictly
            heads = 12 if embed_dim == 768 else 8
            self.proto_fg_attnetion = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=heads, batch_first=True)
            self.proto_bg_attnetion = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=heads, batch_first=True)

            def proj():
                return nn.Sequential(nn.Conv2d(embed_dim, 256, 1), nn.ReLU(inplace=True), nn.Conv2d(256, 128, 1),
                                     nn.ReLU(inplace=True), nn.Conv2d(128, 1, 1))
            self.fg_mask_projection = proj()
            self.bg_mask_projection = proj()
        # An empty 'gridconv' set: the reference prints and raises (RuntimeError from F.conv2d, :193-194 / :68).  Knowing
        # that inside forward() costs a device->host read per call, so by default forward() does NOT synchronise: the
        # scores of such a call are NaN, the set's status word stays on the device in `last_status`, and
        # raise_if_empty() (one 4-byte read, whenever the caller likes -- e.g. once per volume) turns it into the
        # reference's error.  check_empty = True restores the reference's raise-inside-forward behaviour.
        self.check_empty = False
        self.last_status = None
        self.match_algo = 0

    def raise_if_empty(self):
        """The reference's error for the LAST forward() call, on demand (see __init__)."""
        if self.last_status is not None and int(self.last_status[0].item()) & _lib.SET_EMPTY:
            print("failed to find prototypes")      # :193-194, then F.conv2d raises on a [0,C,1,1] weight
            raise RuntimeError("no prototypes survived the threshold: conv2d with a weight of size [0, C, 1, 1]")

    def forward(self, qry, sup_x, sup_y, mode, thresh, isval=False, val_wsize=None, vis_sim=False,
                get_prototypes=False, **kwargs):
        """
        qry:    [way(1), nb(1), nc, h, w] or [nb, nc, h, w]
        sup_x:  [way(1), shot, nb(1), nc, h, w]
        sup_y:  [way(1), shot, nb(1), h, w]
        returns (pred_grid [nb,1,h,w], [debug_assign [nb,h,w]], vis_dict, proto_grid)  -- models/alpmodule.py:161-198
        """
        if mode not in VALID_MODES:
            raise ValueError(f"Invalid mode: {mode}. Expected 'mask', 'gridconv', or 'gridconv+'.")
        if torch.is_grad_enabled() and any(t.requires_grad for t in (qry, sup_x, sup_y)):
            raise RuntimeError("protosam_b200.MultiProtoAsConv is inference-only: call it under torch.no_grad() "
                               "(training / alignLoss use the reference module)")
        qry = qry.squeeze(1)                        # :178
        sup_x = sup_x.squeeze(0).squeeze(1)         # :179  [shot, nc, h, w]
        sup_y = sup_y.squeeze(0)                    # :180
        if val_wsize is None:                       # :187-190
            val_wsize = self.avg_pool_op.kernel_size
            if isinstance(val_wsize, (tuple, list)):
                val_wsize = val_wsize[0]
        S, C, h, w = sup_x.shape
        sup_y = sup_y.reshape(S, 1, h, w)           # :191
        ksize = (val_wsize, val_wsize) if isval else tuple(self.kernel_size)
        ops._need_cuda(qry, sup_x, sup_y)

        protos = ops.alp_prototypes(sup_x, sup_y.reshape(1, S, h, w), [mode], ksize, thresh)
        self.last_status = protos["status"]
        if mode == "gridconv" and self.check_empty:
            self.raise_if_empty()

        Q = qry.shape[0]
        q = qry.permute(0, 2, 3, 1)                 # channels-last view; a no-op for DINOv2 tokens
        if not q.is_contiguous():
            q = q.contiguous()
        scores, assign, sims = ops.alp_match(q.view(Q, h * w, C), protos, want_assign=True, want_sims=vis_sim,
                                             algo=self.match_algo)
        pred_grid = scores[:, 0].view(Q, 1, h, w)
        debug_assign = assign[:, 0].view(Q, h, w)
        vis_dict = {"proto_assign": debug_assign}
        if mode == "mask":
            if vis_sim:
                vis_dict["raw_local_sims"] = debug_assign
            proto_grid = sup_y.clone().detach()     # :104
        else:
            if vis_sim:
                P = int(protos["counts"][0].item())
                vis_dict["raw_local_sims"] = sims[:, 0, :P].reshape(Q, P, h, w)
            gh, gw = protos["gh"], protos["gw"]
            proto_grid = ops.alp_proto_grid(protos["pooled"][0], S, gh, gw, int(val_wsize), thresh, mode)
        return pred_grid, [debug_assign], vis_dict, proto_grid
