"""Drop-in for `infer_module.dynamic_infer_module` (reference infer_module/dynamic_infer_module.py).

Same class names, constructor arguments, parameter names (state_dict keys) and return convention
(`forward(x[B,T,N,C]) -> (y[B,T,N,C], ft_infer_MAD)`), but forward is ONE fused CUDA kernel per sampling
ratio (csrc/head.cu:dynamic_infer_kernel: affinity convs -> softmax over the kt x kn neighbourhood ->
dynamic-walk sampling -> aggregation) followed by the hidden_weight GEMM — instead of ~45 torch ops that
materialise 4 x [B,T,N,k²,C] gathers.

`ft_infer_MAD` ([B,T,N,k²,C] in the reference, :260) has no consumer anywhere in the reference
(SURVEY.md §2 #16), so an empty placeholder tensor is returned in its place.

Autograd: with grad enabled and anything requiring a gradient (the input or a parameter) each class runs as ONE
autograd node (`_DpiFn`): forward = the same kernels with the intermediates kept, backward = the CUDA backward kernels
of csrc/head_bwd.cu (din_dynamic_infer_bwd_f32: gradient to x through the gathers and both affinity convolutions, to the
offsets through the bilinear weights only -- the floor is detached, dynamic_infer_module.py:208 -- and to the relation
logits through the softmax), so the modules train stand-alone exactly as they do inside infer_model's models.
"""
import torch
import torch.nn as nn

from din_b200 import ops
from din_b200 import train as _train
from din_b200.engine import DPIWeights


def _check_input(x):
    if not x.is_cuda:
        raise RuntimeError("Dynamic inference: CUDA tensors required (no CPU fallback on the DIN hot path)")


def _wants_grad(module, x):
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


class _DpiFn(torch.autograd.Function):
    """One autograd node for a stand-alone Dynamic_Person_Inference / Multi_ / Hierarchical_ module.

    chain: [(DPIWeights, parameter-name prefix)].  mode 'sum': every element of the chain sees the same input and the
    outputs are summed (Multi_Dynamic_Inference :436-443; a chain of one = a bare Dynamic_Person_Inference).
    mode 'hier': DPI_1 -> hier_LN -> ReLU -> dropout(0.5, training only) -> DPI_2 (:491-498 with patch H)."""

    @staticmethod
    def forward(ctx, module, mode, chain, x, n_valid, training, names, *params):
        x = x.detach().contiguous().float()
        tape = []
        with torch.no_grad():
            if mode == "sum":
                y = None
                for i, (dpi, _) in enumerate(chain):
                    tmp = _train._dpi_forward(dpi, x, n_valid, tape)
                    y = ops.linear_f32(tmp, dpi.hidden, None, out=y, accumulate=i > 0)
            else:
                (dpi1, _), (dpi2, _) = chain
                B = x.shape[0]
                g = x.numel() // B
                ln = module.hier_LN
                gamma, beta = ln.weight.detach().contiguous(), ln.bias.detach().contiguous()
                tmp1 = _train._dpi_forward(dpi1, x, None, tape)
                y1 = ops.linear_f32(tmp1, dpi1.hidden, None)
                y1n = ops.group_layernorm(y1, gamma, beta, n_outer=B, outer_stride=g, cols=g, relu=True, eps=ln.eps)
                hmask = _train._dropout_mask(y1n.shape, 0.5, x.device, training)
                y1d = ops.scale_mask(y1n, hmask, 2.0) if hmask is not None else y1n
                tmp2 = _train._dpi_forward(dpi2, y1d, None, tape)
                y = ops.linear_f32(tmp2, dpi2.hidden, None)
                ctx.hier = (y1, hmask, gamma, beta, ln.eps, B, g)
        ctx.mode, ctx.chain, ctx.tape, ctx.n_valid, ctx.names = mode, chain, tape, n_valid, names
        ctx.shapes = [tuple(p.shape) for p in params]
        return y

    @staticmethod
    def backward(ctx, dy):
        grads = {}
        dy = dy.contiguous().float()
        with torch.no_grad():
            if ctx.mode == "sum":
                dx = torch.zeros_like(dy)
                for (dpi, prefix), (_, _, xi, tmp) in zip(ctx.chain, ctx.tape):
                    _train._dpi_backward(dpi, prefix, xi, tmp, dy, dx, ctx.n_valid, grads)
            else:
                (dpi1, pre1), (dpi2, pre2) = ctx.chain
                (_, _, x1, tmp1), (_, _, x2, tmp2) = ctx.tape
                y1, hmask, gamma, beta, eps, B, g = ctx.hier
                dy1d = torch.zeros_like(x2)
                _train._dpi_backward(dpi2, pre2, x2, tmp2, dy, dy1d, None, grads)
                dy1n = ops.scale_mask(dy1d, hmask, 2.0) if hmask is not None else dy1d
                dy1, dgam, dbet = ops.group_layernorm_bwd(y1, gamma, beta, dy1n, n_outer=B, outer_stride=g, cols=g,
                                                          relu=True, eps=eps)
                grads["hier_LN.weight"], grads["hier_LN.bias"] = dgam.view_as(gamma), dbet.view_as(beta)
                dx = torch.zeros_like(x1)
                _train._dpi_backward(dpi1, pre1, x1, tmp1, dy1, dx, None, grads)
        ctx.tape = None
        out = [grads[n].reshape(shp) if (need and n in grads) else None
               for n, shp, need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[7:])]
        return (None, None, None, dx if ctx.needs_input_grad[3] else None, None, None, None) + tuple(out)


def _run_with_grad(module, mode, chain, x, n_valid=None):
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    return _DpiFn.apply(module, mode, chain, x, n_valid, module.training, tuple(n for n, _ in named),
                        *[p for _, p in named])


class Dynamic_Person_Inference(nn.Module):
    """reference :14-404.  Only the configuration every stage-2 script uses is implemented:
    dynamic_sampling=True, parallel_inference=False, stride=1, group=1."""

    def __init__(self, in_dim, person_mat_shape, stride=1, kernel_size=(3, 3), dynamic_sampling=False,
                 sampling_ratio=[1], group=1, scale_factor=False, beta_factor=False, parallel_inference=False,
                 cfg=None):
        super().__init__()
        if not dynamic_sampling:
            raise NotImplementedError("dynamic_sampling=False is a dead path in the reference (forward raises "
                                      "UnboundLocalError at :151)")
        if parallel_inference:
            raise NotImplementedError("parallel_inference=True is outside the hot-path scope (no config uses it)")
        if stride != 1 or group != 1:
            raise NotImplementedError("stride/group other than 1 are never used by the reference scripts")
        self.T, self.N = person_mat_shape
        self.stride, self.kernel_size = stride, tuple(kernel_size)
        self.dynamic_sampling, self.sampling_ratio = dynamic_sampling, list(sampling_ratio)
        self.scale_factor, self.beta_factor = scale_factor, beta_factor
        self.max_ratio = self.sampling_ratio[-1]
        self.parallel_inference, self.cfg = parallel_inference, cfg
        kt, kn = self.kernel_size

        self.hidden_weight = nn.Linear(in_dim, in_dim, bias=False)
        if beta_factor:
            self.beta = nn.Parameter(torch.ones(len(self.sampling_ratio)), requires_grad=True)
        self.zero_padding = nn.ModuleDict()
        self.p_conv = nn.ModuleDict()
        if scale_factor:
            self.scale_conv = nn.ModuleDict()
        for r in self.sampling_ratio:
            pad_tb, pad_lr = (kt - 1) // 2 * r, (kn - 1) // 2 * r
            self.zero_padding[str(r)] = nn.ZeroPad2d((pad_lr, pad_lr, pad_tb, pad_tb))
            # parameter containers only: the convolution itself runs inside the fused kernel
            pc = nn.Conv2d(in_dim, 2 * kt * kn, self.kernel_size, dilation=r, stride=stride,
                           padding=(pad_tb, pad_lr), groups=group, bias=True)
            nn.init.zeros_(pc.weight)
            nn.init.zeros_(pc.bias)
            self.p_conv[str(r)] = pc
            if scale_factor:
                sc = nn.Conv2d(in_dim, kt * kn, self.kernel_size, dilation=r, stride=stride,
                               padding=(pad_tb, pad_lr), groups=group, bias=True)
                nn.init.zeros_(sc.weight)
                nn.init.zeros_(sc.bias)
                self.scale_conv[str(r)] = sc
        nn.init.kaiming_normal_(self.hidden_weight.weight)
        self._packed, self._packed_key = None, None

    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._packed = DPIWeights(sd, "", self.kernel_size, self.sampling_ratio, self.scale_factor,
                                      self.beta_factor)
            self._packed_key = key
        return self._packed

    def forward(self, person_features, n_valid=None):
        _check_input(person_features)
        if _wants_grad(self, person_features):
            y = _run_with_grad(self, "sum", [(self._weights(), "")], person_features, n_valid)
            return y, y.new_empty(0)
        x = person_features.detach().contiguous().float()
        with torch.no_grad():
            y = self._weights()(x, n_valid=n_valid)
        return y, x.new_empty(0)


class Multi_Dynamic_Inference(nn.Module):
    """reference :407-443 — num_DIM modules on the same input, outputs summed."""

    def __init__(self, in_dim, person_mat_shape, stride=1, kernel_size=[(3, 3)], dynamic_sampling=False,
                 sampling_ratio=[1], group=1, scale_factor=False, beta_factor=False, parallel_inference=False,
                 num_DIM=1, cfg=None):
        super().__init__()
        self.DIMlist = nn.ModuleList([
            Dynamic_Person_Inference(in_dim=in_dim, person_mat_shape=person_mat_shape, stride=stride,
                                     kernel_size=kernel_size[i], dynamic_sampling=dynamic_sampling,
                                     sampling_ratio=sampling_ratio, group=group, scale_factor=scale_factor,
                                     beta_factor=beta_factor, parallel_inference=parallel_inference, cfg=cfg)
            for i in range(num_DIM)])

    def forward(self, person_features):
        _check_input(person_features)
        if _wants_grad(self, person_features):
            chain = [(dim._weights(), f"DIMlist.{i}.") for i, dim in enumerate(self.DIMlist)]
            y = _run_with_grad(self, "sum", chain, person_features)
            return y, y.new_empty(0)
        x = person_features.detach().contiguous().float()
        out = None
        with torch.no_grad():
            for i, dim in enumerate(self.DIMlist):
                out = dim._weights()(x, out=out, accumulate=i > 0)
        return out, x.new_empty(0)


class Hierarchical_Dynamic_Inference(nn.Module):
    """reference :446-498 with its defects repaired (SURVEY.md §8c bug H): DPI_1 -> LayerNorm -> ReLU ->
    dropout (train only) -> DPI_2, returning (out, mad).  hier_LN keeps the reference's hard-coded shape
    person_mat_shape + (1024,)."""

    def __init__(self, in_dim, person_mat_shape, stride=1, kernel_size=[(3, 3)], dynamic_sampling=False,
                 sampling_ratio=[1], group=1, scale_factor=False, beta_factor=False, parallel_inference=False,
                 cfg=None):
        super().__init__()
        assert len(kernel_size) == 2
        kw = dict(in_dim=in_dim, person_mat_shape=person_mat_shape, stride=stride,
                  dynamic_sampling=dynamic_sampling, sampling_ratio=sampling_ratio, group=group,
                  scale_factor=scale_factor, beta_factor=beta_factor, parallel_inference=parallel_inference,
                  cfg=cfg)
        self.DPI_1 = Dynamic_Person_Inference(kernel_size=kernel_size[0], **kw)
        self.hier_LN = nn.LayerNorm(tuple(person_mat_shape) + (1024,))
        self.dropout = nn.Dropout(0.3)
        self.DPI_2 = Dynamic_Person_Inference(kernel_size=kernel_size[1], **kw)

    def forward(self, person_features):
        _check_input(person_features)
        if _wants_grad(self, person_features) or self.training:
            chain = [(self.DPI_1._weights(), "DPI_1."), (self.DPI_2._weights(), "DPI_2.")]
            y = _run_with_grad(self, "hier", chain, person_features)
            return y, y.new_empty(0)
        x = person_features.detach().contiguous().float()
        with torch.no_grad():
            y1 = self.DPI_1._weights()(x)
            B = x.shape[0]
            g = y1.numel() // B
            y1 = ops.group_layernorm(y1, self.hier_LN.weight.detach().contiguous(),
                                     self.hier_LN.bias.detach().contiguous(), n_outer=B, outer_stride=g, cols=g,
                                     relu=True, eps=self.hier_LN.eps)
            y2 = self.DPI_2._weights()(y1)
        return y2, x.new_empty(0)
