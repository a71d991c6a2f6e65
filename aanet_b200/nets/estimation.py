"""Soft-argmin disparity regression -- drop-in for the reference's nets/estimation.py."""
import torch.nn as nn

from .. import ops


class DisparityEstimation(nn.Module):
    """`DisparityEstimation(max_disp, match_similarity=True)`; forward([B,D,H,W]) -> [B,H,W].

    As in the reference (estimation.py:21-25) the disparity candidates are 0..D-1 of the tensor that
    is passed in, whatever `max_disp` says."""

    def __init__(self, max_disp, match_similarity=True):
        super().__init__()
        self.max_disp = max_disp
        self.match_similarity = match_similarity

    def forward(self, cost_volume):
        assert cost_volume.dim() == 4
        return ops.soft_argmin(cost_volume, self.match_similarity)
