"""Drop-in for `roi_align.roi_align.RoIAlign` (longcw/RoIAlign.pytorch; reference infer_model.py:3,48,178).

Same constructor and call signature; the arithmetic is csrc/head.cu:roi_align_kernel through the C ABI
(din_roi_align_nhwc_f16).  The fused models in infer_model.py call the kernel on the NHWC fp16 map
directly; this module is the stand-alone surface and converts layouts around the same kernel.
"""
import torch
import torch.nn as nn

from din_b200 import ops


class RoIAlign(nn.Module):
    def __init__(self, crop_height, crop_width, extrapolation_value=0, transform_fpcoor=True):
        super().__init__()
        if extrapolation_value != 0 or not transform_fpcoor:
            raise NotImplementedError("only extrapolation_value=0, transform_fpcoor=True (the values the "
                                      "reference uses) are implemented")
        self.crop_height, self.crop_width = crop_height, crop_width

    def forward(self, featuremap, boxes, box_ind):
        """featuremap [B,D,H,W], boxes [M,4] (x1,y1,x2,y2), box_ind [M] int -> [M,D,crop_h,crop_w]."""
        if not featuremap.is_cuda:
            raise RuntimeError("RoIAlign: CUDA tensors required (no CPU fallback on the DIN hot path)")
        b, d, h, w = featuremap.shape
        d8 = (d + 7) // 8 * 8
        fm = torch.zeros((b, h, w, d8), dtype=torch.float16, device=featuremap.device)
        fm[..., :d] = featuremap.permute(0, 2, 3, 1)
        out = ops.roi_align_nhwc(fm, boxes.detach().float().contiguous(), box_ind.detach().int().contiguous(),
                                 self.crop_height, self.crop_width, d=d8)
        m = boxes.shape[0]
        out = out.view(m, self.crop_height, self.crop_width, d8)[..., :d]
        return out.permute(0, 3, 1, 2).to(featuremap.dtype).contiguous()
