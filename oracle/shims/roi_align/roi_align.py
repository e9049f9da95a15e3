"""Import shim for the un-vendored `longcw/RoIAlign.pytorch` extension (reference Dockerfile:6-8,
infer_model.py:3).  Lets the reference's own classes be imported in this container; the arithmetic is
the oracle's restatement of that package's published algorithm (oracle/din_oracle.py:roi_align_longcw)."""
import os
import sys

import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from din_oracle import roi_align_longcw  # noqa: E402


class RoIAlign(nn.Module):
    def __init__(self, crop_height, crop_width, extrapolation_value=0, transform_fpcoor=True):
        super().__init__()
        assert transform_fpcoor
        self.crop_height, self.crop_width, self.extrapolation_value = crop_height, crop_width, extrapolation_value

    def forward(self, featuremap, boxes, box_ind):
        return roi_align_longcw(featuremap, boxes, box_ind, self.crop_height, self.crop_width,
                                float(self.extrapolation_value))
