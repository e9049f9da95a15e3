"""train.py — the stage-2 (and stage-1) training step on one B200.

`forward_train` is DinEngine's forward with dropout applied and the intermediates the backward needs kept on a
tape; `backward_head` turns d(loss)/d(logits) into gradients for every parameter after the backbone and -- with
cfg.train_backbone and the VGG-16 / ResNet-18 backbone -- for the backbone too (`backward_backbone`), under the
reference's parameter names and layouts (so `optimizer.step()` in train_net_dynamic.py:220-224 works on the
drop-in model unchanged).  With the backbone frozen (config.py:39 `train_backbone = False`) gradients stop at the
feature map.  A data-parallel run receives the gradients in two groups as they become final (`sink`, see
parallel.BucketedGradientReducer).  Host-side orchestration only: every number is produced by a kernel of
libdin_sm100.so (ops.py); torch is used for buffers, views and layout permutes.

Reference graph (infer_model.py:141-234 / 1226-1319, infer_module/dynamic_infer_module.py:121-151, 407-498):
  crops -> fc_emb_1 -> nl_emb_1 -> ReLU -> [point_conv -> point_ln -> ReLU] = x
  g = sum_i hidden_i( mean_r | sum_r beta_r  DIN_{i,r}(x) )          (Multi)   or   DPI_2(drop(relu(LN(DPI_1(x)))))
  s = relu(dpi_nl(g + x)) (vgg16/inv3)  |  relu(dpi_nl(g)) + x (res18)  |  Collective: LN([T,C]) per actor of g + x
  logits = mean_t fc_activities(max_n dropout(s))
"""
from __future__ import annotations

import torch

from . import ops


def _dropout_mask(shape, p, device, training):
    """0/1 byte mask from torch's CUDA generator (reproducible with torch.manual_seed), or None when inactive."""
    if not training or p <= 0.0:
        return None
    return (torch.rand(shape, device=device) >= p).to(torch.uint8)


def _dpi_forward(dpi, x, n_valid, tape):
    """DPIWeights.__call__ keeping `tmp` (the mixed ratio outputs, input of hidden_weight)."""
    tmp = torch.zeros_like(x) if n_valid is not None else torch.empty_like(x)
    for i, r in enumerate(dpi.ratios):
        w_tap, b_cat = dpi.taps[r]
        if dpi.beta_factor:
            ops.dynamic_infer(x, w_tap, b_cat, dpi.kernel, r, scale_factor=dpi.scale_factor, out=tmp,
                              coef_ptr=dpi.beta.data_ptr() + 4 * i, accumulate=i > 0, n_valid=n_valid)
        else:
            ops.dynamic_infer(x, w_tap, b_cat, dpi.kernel, r, scale_factor=dpi.scale_factor, out=tmp,
                              coef=1.0 / len(dpi.ratios), accumulate=i > 0, n_valid=n_valid)
    tape.append(("dpi", dpi, x, tmp))
    return tmp


def _dpi_backward(dpi, prefix, x, tmp, dg, dx, n_valid, grads):
    """g = tmp . hidden^T ; tmp = sum_r coef_r DIN_r(x).  dx += d/dx; parameter gradients -> grads[prefix + ...]."""
    C = x.shape[-1]
    M = x.numel() // C
    dtmp, dwh, _ = ops.linear_bwd(tmp.view(M, C), dpi.hidden, dg.view(M, C), has_bias=False)
    grads[prefix + "hidden_weight.weight"] = dwh
    kt, kn = dpi.kernel
    k2 = kt * kn
    dbeta = []
    for i, r in enumerate(dpi.ratios):
        w_tap, b_cat = dpi.taps[r]
        kw = dict(coef_ptr=dpi.beta.data_ptr() + 4 * i, want_dcoef=True) if dpi.beta_factor else \
            dict(coef=1.0 / len(dpi.ratios))
        dw, db, dcoef = ops.dynamic_infer_bwd(x, w_tap, b_cat, dtmp.view_as(x), dx, dpi.kernel, r,
                                              scale_factor=dpi.scale_factor, n_valid=n_valid, **kw)
        # packed [tap][o][c] -> OIHW [o, c, kt, kn] (layout only)
        w_oihw = dw.permute(1, 2, 0).reshape(dw.shape[1], C, kt, kn)
        grads[f"{prefix}p_conv.{r}.weight"] = w_oihw[:2 * k2].contiguous()
        grads[f"{prefix}p_conv.{r}.bias"] = db[:2 * k2].contiguous()
        if dpi.scale_factor:
            grads[f"{prefix}scale_conv.{r}.weight"] = w_oihw[2 * k2:].contiguous()
            grads[f"{prefix}scale_conv.{r}.bias"] = db[2 * k2:].contiguous()
        if dpi.beta_factor:
            dbeta.append(dcoef)
    if dpi.beta_factor:
        grads[prefix + "beta"] = torch.cat(dbeta)


TCE_PREFIX = "multilayer_head_embfeature_context_encoding.CET."


def _tce_forward(tce, x, fm, n, tape, training, p):
    """TCEWeights.__call__ with the intermediates kept and the module's two dropouts applied (TCE_STBiP_module.py:280,
    :241 -- nn.Dropout(context_dropout_ratio) on the attended features and inside the FFN)."""
    from .engine import TCE_DIM as D, TCE_HEADS as H
    B, T, N, C = x.shape
    M = B * T * N
    F_, oh, ow, _ = fm.shape
    dev = x.device
    xf = x.reshape(M, C)
    q = torch.empty((H, M, D), dtype=torch.float32, device=dev)
    for h in range(H):
        ops.linear_f32(xf, *tce.q[h], out=q[h])
    img = tce.conv(fm, out_f32=True).view(F_, oh * ow, H * D)
    ctx = ops.context_attention(q, img, tce.posbias, n)
    out = torch.empty((B, T, N, C + H * D), dtype=torch.float32, device=dev)
    o2 = out.view(M, C + H * D)
    o2[:, :C].copy_(xf)
    heads = []
    scale = 1.0 / (1.0 - p) if p < 1.0 else 0.0
    for h in range(H):
        m1 = _dropout_mask(ctx[h].shape, p, dev, training)
        ctx_d = ops.scale_mask(ctx[h], m1, scale) if m1 is not None else ctx[h]
        c1 = ops.group_layernorm(ctx_d, *tce.ln1[h], n_outer=M, outer_stride=D, cols=D, pre=q[h])
        a = ops.linear_f32(c1, *tce.ffn0[h], relu=True)
        m2 = _dropout_mask(a.shape, p, dev, training)
        a_d = ops.scale_mask(a, m2, scale) if m2 is not None else a
        f = ops.linear_f32(a_d, *tce.ffn3[h])
        o = ops.group_layernorm(f, *tce.ln2[h], n_outer=M, outer_stride=D, cols=D, pre=c1)
        o2[:, C + h * D:C + (h + 1) * D].copy_(o)
        heads.append((m1, ctx_d, c1, a, m2, a_d, f))
    tape["tce"] = {"xf": xf, "q": q, "img": img, "fm": fm, "heads": heads, "scale": scale, "C": C}
    return out


def _tce_backward(eng, tape, dx_full, grads, need_dfm):
    """dx_full [M, C + 512]: gradient of the DIN input.  -> gradient of the person features [M, C]; parameter gradients
    under the reference's names; tape['dfm_init'] = the context encoder's share of the feature-map gradient (fp32)."""
    from .engine import TCE_DIM as D, TCE_HEADS as H
    tce, rec = eng.tce, tape["tce"]
    C, scale = rec["C"], rec["scale"]
    xf, q, img = rec["xf"], rec["q"], rec["img"]
    M = xf.shape[0]
    dxp = dx_full[:, :C].contiguous()
    dctx = torch.empty_like(q)
    dq_res = torch.empty_like(q)
    for h in range(H):
        m1, ctx_d, c1, a, m2, a_d, f = rec["heads"][h]
        pre = f"{TCE_PREFIX}{h}."
        do = dx_full[:, C + h * D:C + (h + 1) * D].contiguous()
        # layernorm2(f + c1): one gradient for both
        df, dg2, db2 = ops.group_layernorm_bwd(f, *tce.ln2[h], do, n_outer=M, outer_stride=D, cols=D, pre=c1)
        grads[pre + "layernorm2.weight"], grads[pre + "layernorm2.bias"] = dg2, db2
        da_d, dw3, db3 = ops.linear_bwd(a_d, tce.ffn3[h][0], df)
        grads[pre + "FFN.3.weight"], grads[pre + "FFN.3.bias"] = dw3, db3
        da = ops.scale_mask(da_d, m2, scale) if m2 is not None else da_d
        dpre = ops.relu_bwd_f32(a, da)
        _, dw0, db0 = ops.linear_bwd(c1, tce.ffn0[h][0], dpre, dx_out=df, dx_accumulate=True)     # df += d(c1) via the FFN
        grads[pre + "FFN.0.weight"], grads[pre + "FFN.0.bias"] = dw0, db0
        # layernorm1(ctx_d + q): one gradient for the attended features and for q's residual
        du, dg1, db1 = ops.group_layernorm_bwd(ctx_d, *tce.ln1[h], df, n_outer=M, outer_stride=D, cols=D, pre=q[h],
                                               dx_out=dq_res[h])
        grads[pre + "layernorm1.weight"], grads[pre + "layernorm1.bias"] = dg1, db1
        if m1 is not None:
            ops.scale_mask(du, m1, scale, out=dctx[h])
        else:
            dctx[h].copy_(du)
    dq, dimg = ops.context_attention_bwd(q, img, tce.posbias, dctx, dq_add=dq_res)
    for h in range(H):
        pre = f"{TCE_PREFIX}{h}."
        _, dwq, dbq = ops.linear_bwd(xf, tce.q[h][0], dq[h], dx_out=dxp, dx_accumulate=True)
        grads[pre + "emb_roi.weight"], grads[pre + "emb_roi.bias"] = dwq, dbq
    # downsample2 of the four heads = one 512 -> 512 GEMM over the map (the bias entered through posbias)
    fm = rec["fm"]
    rows = fm.shape[0] * fm.shape[1] * fm.shape[2]
    dfm, dwd, dbd = ops.linear_bwd(fm.view(rows, fm.shape[3]), tce.w_all, dimg.view(rows, H * D), need_dx=need_dfm)
    # downsample2 runs on context = feature map + position embedding (infer_model.py:413): the embedding's share of dW is
    # (sum over frames of dimg)^T . pos
    F_, P = fm.shape[0], fm.shape[1] * fm.shape[2]
    dsum = ops.mean_axis(dimg.view(1, F_, P * H * D), 1)
    ops.gemm_f32(dsum, tce.pos, m=H * D, n=fm.shape[3], k=P, a_strides=(1, H * D), b_strides=(fm.shape[3], 1), out=dwd,
                 alpha=float(F_), accumulate=True)
    for h in range(H):
        pre = f"{TCE_PREFIX}{h}."
        grads[pre + "downsample2.weight"] = dwd[h * D:(h + 1) * D].reshape(D, fm.shape[3], 1, 1).contiguous()
        grads[pre + "downsample2.bias"] = dbd[h * D:(h + 1) * D].contiguous()
    tape["dfm_init"] = dfm.view(fm.shape[0], fm.shape[1], fm.shape[2], fm.shape[3]) if need_dfm else None
    return dxp


def forward_train(eng, images, boxes, bboxes_num=None, training=True, train_backbone=False, tce_dropout=None):
    """-> (logits [B, A], tape).  Same kernels as DinEngine.forward_*, plus dropout and saved intermediates.
    train_backbone: additionally keep every backbone activation (VGG-16) for backward_backbone."""
    cfg = eng.cfg
    B, T = images.shape[:2]
    N, C = eng.N, eng.C
    M = B * T * N
    dev = eng.device
    tape = {"B": B, "T": T, "dpi": [], "train_backbone": train_backbone}
    with torch.no_grad():
        if train_backbone:
            fm, tape["bb_chunks"] = eng.features_train(eng._flat_frames(images))
            tape["boxes"] = boxes.reshape(M, 4).contiguous().float()
            tape["fm_shape"] = tuple(fm.shape)
        else:
            fm = eng.features(eng._flat_frames(images))
        if eng.dataset == "volleyball":
            OH, OW = cfg.out_size
            assert fm.shape[1:3] == (OH, OW), (tuple(fm.shape), cfg.out_size)
        # ---- embed (DinEngine.embed with the intermediates kept)
        crops = ops.roi_align_nhwc(fm, boxes.reshape(M, 4).contiguous().float(), eng._box_idx(B * T, N), eng.K, eng.K,
                                   d=eng.D_stride)
        emb = eng.fc_emb(crops.view(1, 1, M, eng.K * eng.K * eng.D_stride), out_f32=True).view(M, eng.NFB)
        x0 = ops.group_layernorm(emb, *eng.nl_emb, n_outer=M, outer_stride=eng.NFB, cols=eng.NFB, relu=True)
        tape.update(crops=crops.view(M, -1), emb=emb, x0=x0)
        g_sz = T * N * C
        if cfg.lite_dim:
            ylite = ops.linear_f32(x0, eng.point_w, eng.point_b)
            x = ops.group_layernorm(ylite, *eng.point_ln, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True)
            tape["ylite"] = ylite
        else:
            x = x0
        if eng.tce is not None:                       # Dynamic_TCE_volleyball: 4 x 128 context features join the DIN input
            p_tce = float(getattr(eng, "tce_dropout", 0.1) if tce_dropout is None else tce_dropout)
            x = _tce_forward(eng.tce, x.view(B, T, N, eng.C_embed), fm, N, tape, training, p_tce)
        x = x.view(B, T, N, C)
        tape["x"] = x
        n_valid = None
        if eng.dataset == "collective":
            n_valid = bboxes_num.reshape(B, T)[:, 0].to(torch.int32).contiguous()
        tape["n_valid"] = n_valid
        # ---- dynamic inference
        if eng.hier:
            tmp1 = _dpi_forward(eng.dpis[0], x, None, tape["dpi"])
            y1 = ops.linear_f32(tmp1, eng.dpis[0].hidden, None)
            y1n = ops.group_layernorm(y1, *eng.hier_ln, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True)
            # F.dropout(p=0.5) (dynamic_infer_module.py:495; honours train/eval under oracle patch H)
            hmask = _dropout_mask(y1n.shape, 0.5, dev, training)
            y1d = ops.scale_mask(y1n, hmask, 2.0) if hmask is not None else y1n
            tape.update(y1=y1, hmask=hmask, y1d=y1d)
            tmp2 = _dpi_forward(eng.dpis[1], y1d, None, tape["dpi"])
            g = ops.linear_f32(tmp2, eng.dpis[1].hidden, None)
        else:
            g = None
            for i, dpi in enumerate(eng.dpis):
                tmp = _dpi_forward(dpi, x, n_valid, tape["dpi"])
                g = ops.linear_f32(tmp, dpi.hidden, None, out=g, accumulate=i > 0)
        tape["g"] = g
        # ---- fusion + LayerNorm + ReLU
        if eng.dataset == "collective":
            s = torch.zeros_like(g)
            ops.group_layernorm(g, *eng.dpi_nl, n_outer=B, n_inner=N, outer_stride=g_sz, inner_stride=C, rows=T,
                                row_stride=N * C, cols=C, relu=True, pre=x, n_valid=n_valid, out=s)
        elif eng.backbone_name == "res18":
            s = ops.group_layernorm(g, *eng.dpi_nl, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True, post=x)
        else:
            s = ops.group_layernorm(g, *eng.dpi_nl, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True, pre=x)
        # ---- dropout_global (infer_model.py:209,216 / :1303) and read-out
        p = float(cfg.train_dropout_prob)
        mask = _dropout_mask(s.shape, p, dev, training)
        s_d = ops.scale_mask(s, mask, 1.0 / (1.0 - p)) if mask is not None else s
        tape.update(mask=mask, s_d=s_d, p=p)
        logits = ops.readout(s_d, *eng.fc_act, n_valid=n_valid)
    return logits, tape


def backward_head(eng, tape, dlogits, sink=None):
    """d(loss)/d(logits) [B, A] -> {reference parameter name: gradient} for every parameter after the backbone (and,
    with tape['train_backbone'], the backbone's).  sink(stage, grads) is called as soon as a group of gradients is
    final -- "head" before the backbone's backward starts, "backbone" at the end -- so that a data-parallel run can
    exchange the head's gradients while the backbone's backward still runs (parallel.BucketedGradientReducer)."""
    cfg = eng.cfg
    B, T = tape["B"], tape["T"]
    N, C, NFB = eng.N, eng.C, eng.NFB
    M = B * T * N
    g_sz = T * N * C
    n_valid = tape["n_valid"]
    x, g = tape["x"], tape["g"]
    grads = {}
    with torch.no_grad():
        # ---- read-out and dropout
        ds_d, dwa, dba = ops.readout_bwd(tape["s_d"], eng.fc_act[0], dlogits.contiguous().float(), n_valid=n_valid)
        grads["fc_activities.weight"], grads["fc_activities.bias"] = dwa, dba
        ds = ops.scale_mask(ds_d, tape["mask"], 1.0 / (1.0 - tape["p"])) if tape["mask"] is not None else ds_d
        # ---- fusion LayerNorm: dg, and the gradient x receives directly (dx)
        if eng.dataset == "collective":
            dg, dgam, dbet = ops.group_layernorm_bwd(g, *eng.dpi_nl, ds, n_outer=B, n_inner=N, outer_stride=g_sz,
                                                     inner_stride=C, rows=T, row_stride=N * C, cols=C, relu=True,
                                                     pre=x, n_valid=n_valid)
            dx = dg.clone()
        elif eng.backbone_name == "res18":
            dg, dgam, dbet = ops.group_layernorm_bwd(g, *eng.dpi_nl, ds, n_outer=B, outer_stride=g_sz, cols=g_sz,
                                                     relu=True)
            dx = ds.clone()                                    # `+ x` after the ReLU (infer_model.py:207)
        else:
            dg, dgam, dbet = ops.group_layernorm_bwd(g, *eng.dpi_nl, ds, n_outer=B, outer_stride=g_sz, cols=g_sz,
                                                     relu=True, pre=x)
            dx = dg.clone()                                    # LN(g + x): x gets the same gradient as g
        grads["dpi_nl.weight"], grads["dpi_nl.bias"] = dgam.view_as(eng.dpi_nl[0]), dbet.view_as(eng.dpi_nl[1])
        # ---- dynamic inference
        if eng.hier:
            (_, dpi1, x1, tmp1), (_, dpi2, x2, tmp2) = tape["dpi"]
            dy1d = torch.zeros_like(x2)
            _dpi_backward(dpi2, "DPI.DPI_2.", x2, tmp2, dg, dy1d, None, grads)
            dy1n = ops.scale_mask(dy1d, tape["hmask"], 2.0) if tape["hmask"] is not None else dy1d
            dy1, dgam, dbet = ops.group_layernorm_bwd(tape["y1"], *eng.hier_ln, dy1n, n_outer=B, outer_stride=g_sz,
                                                      cols=g_sz, relu=True)
            grads["DPI.hier_LN.weight"] = dgam.view_as(eng.hier_ln[0])
            grads["DPI.hier_LN.bias"] = dbet.view_as(eng.hier_ln[1])
            _dpi_backward(dpi1, "DPI.DPI_1.", x1, tmp1, dy1, dx, None, grads)
        else:
            for i, (_, dpi, xi, tmp) in enumerate(tape["dpi"]):
                prefix = "DPI." if eng.dataset == "collective" else f"DPI.DIMlist.{i}."
                _dpi_backward(dpi, prefix, xi, tmp, dg, dx, n_valid, grads)
        # ---- context encoder (Dynamic_TCE_volleyball): the DIN input is [person features | 4 x 128 context features]
        dx = dx.view(M, C)
        if eng.tce is not None:
            dx = _tce_backward(eng, tape, dx, grads, need_dfm=tape["train_backbone"])
            C = eng.C_embed
            g_sz = T * N * C
        if cfg.lite_dim:
            dyl, dgam, dbet = ops.group_layernorm_bwd(tape["ylite"], *eng.point_ln, dx, n_outer=B, outer_stride=g_sz,
                                                      cols=g_sz, relu=True)
            grads["point_ln.weight"] = dgam.view_as(eng.point_ln[0])
            grads["point_ln.bias"] = dbet.view_as(eng.point_ln[1])
            dx0, dwp, dbp = ops.linear_bwd(tape["x0"], eng.point_w, dyl)
            grads["point_conv.weight"], grads["point_conv.bias"] = dwp.view(C, NFB, 1, 1), dbp
        else:
            dx0 = dx
        # ---- nl_emb_1 + fc_emb_1
        demb, dgam, dbet = ops.group_layernorm_bwd(tape["emb"], *eng.nl_emb, dx0, n_outer=M, outer_stride=NFB,
                                                   cols=NFB, relu=True)
        grads["nl_emb_1.weight"], grads["nl_emb_1.bias"] = dgam, dbet
        # dW in the kernel's K order (ky, kx, d_padded) -> the reference's (d, ky, kx) flatten (layout only)
        dcrops, dwk, dbe = ops.linear_bwd(tape["crops"], eng.fc_emb_wk if tape["train_backbone"] else None, demb,
                                          need_dx=tape["train_backbone"])
        KK = eng.K * eng.K
        grads["fc_emb_1.weight"] = dwk.view(NFB, KK, eng.D_stride)[:, :, :eng.D].permute(0, 2, 1).reshape(NFB, -1) \
            .contiguous()
        grads["fc_emb_1.bias"] = dbe
        if sink is not None:
            sink("head", grads)
        if tape["train_backbone"]:
            backward_backbone(eng, tape, dcrops, grads)
            if sink is not None:
                sink("backbone", grads)
    return grads


def backward_backbone(eng, tape, dcrops, grads):
    """d(loss)/d(crops) [M, 25*D_stride] fp32 -> RoIAlign backward -> scaled fp16 -> VGG-16 backward (dgrad on the
    forward tcgen05 kernel, wgrad on conv_wgrad_tcgen05.cu), chunk by chunk; gradients under the reference's
    `backbone.features.N.{weight,bias}` names."""
    B, T = tape["B"], tape["T"]
    N = eng.N
    dfm = tape.get("dfm_init")                                      # the context encoder's share (Dynamic_TCE_volleyball)
    if dfm is None:
        dfm = torch.zeros(tape["fm_shape"], dtype=torch.float32, device=eng.device)
    ops.roi_align_bwd(dcrops, tape["boxes"], eng._box_idx(B * T, N), dfm, eng.K, eng.K, d=eng.D_stride)
    dfm16, scale_ws = ops.grad_to_f16(dfm)
    inv_scale = scale_ws[2:3]
    acc = eng.backbone.new_grads(eng.device)
    for f0, f1, saved in tape["bb_chunks"]:
        eng.backbone.backward(saved, dfm16[f0:f1], inv_scale, acc)
    eng.backbone.export_grads(acc, grads)


# ------------------------------------------------------------------------------------------------------------------
# stage 1: Basenet_volleyball (base_model.py:64-142) -- scripts/train_volleyball_stage1.py trains it with the VGG-16
# backbone, T = 1, cross-entropy on the activities plus weighted cross-entropy on the per-actor actions
# ------------------------------------------------------------------------------------------------------------------
def basenet_forward_train(eng, images, boxes, training=True, train_backbone=True):
    """-> ((actions [B*N, A], activities [B, A2]), tape)."""
    cfg = eng.cfg
    B, T = images.shape[:2]
    N, NFB = eng.N, eng.NFB
    M = B * T * N
    tape = {"B": B, "T": T, "train_backbone": train_backbone}
    with torch.no_grad():
        if train_backbone:
            fm, tape["bb_chunks"] = eng.features_train(eng._flat_frames(images))
            tape["boxes"] = boxes.reshape(M, 4).contiguous().float()
            tape["fm_shape"] = tuple(fm.shape)
        else:
            fm = eng.features(eng._flat_frames(images))
        crops = ops.roi_align_nhwc(fm, boxes.reshape(M, 4).contiguous().float(), eng._box_idx(B * T, N), eng.K, eng.K,
                                   d=eng.D_stride)
        x = eng.fc_emb(crops.view(1, 1, M, eng.K * eng.K * eng.D_stride), out_f32=True).view(M, NFB)   # ReLU fused
        p = float(cfg.train_dropout_prob)
        mask = _dropout_mask(x.shape, p, eng.device, training)                                         # dropout_emb :121
        x_d = ops.scale_mask(x, mask, 1.0 / (1.0 - p)) if mask is not None else x
        tape.update(crops=crops.view(M, -1), x=x, x_d=x_d, mask=mask, p=p)
        actions = ops.linear_f32(x_d, *eng.fc_actions)
        activities = ops.readout(x_d.view(B, T, N, NFB), *eng.fc_act)
        if T != 1:
            actions = ops.mean_axis(actions.view(B, T, N * actions.shape[-1]), 1)
    return (actions.view(B * N, -1), activities), tape


def basenet_backward(eng, tape, dactions, dactivities, sink=None):
    """(d/d actions [B*N, A], d/d activities [B, A2]) -> {reference parameter name: gradient}; `sink` as in
    backward_head."""
    B, T = tape["B"], tape["T"]
    N, NFB = eng.N, eng.NFB
    M = B * T * N
    grads = {}
    with torch.no_grad():
        x_d = tape["x_d"]
        dx, dwa, dba = ops.readout_bwd(x_d.view(B, T, N, NFB), eng.fc_act[0], dactivities.contiguous().float())
        grads["fc_activities.weight"], grads["fc_activities.bias"] = dwa, dba
        A = eng.fc_actions[0].shape[0]
        dact = dactions.contiguous().float().view(B, 1, N, A)
        if T != 1:                                             # mean over frames (:138-139): broadcast / T
            dact = ops.scale_mask(dact.expand(B, T, N, A).contiguous(), None, 1.0 / T)
        _, dw, db = ops.linear_bwd(x_d, eng.fc_actions[0], dact.reshape(M, A), dx_out=dx.view(M, NFB), dx_accumulate=True)
        grads["fc_actions.weight"], grads["fc_actions.bias"] = dw, db
        dx = dx.view(M, NFB)
        if tape["mask"] is not None:
            dx = ops.scale_mask(dx, tape["mask"], 1.0 / (1.0 - tape["p"]))
        demb = ops.relu_bwd_f32(tape["x"], dx)                 # F.relu (:120)
        dcrops, dwk, dbe = ops.linear_bwd(tape["crops"], eng.fc_emb_wk if tape["train_backbone"] else None, demb,
                                          need_dx=tape["train_backbone"])
        KK = eng.K * eng.K
        name = eng.emb_name
        grads[name + ".weight"] = dwk.view(NFB, KK, eng.D_stride)[:, :, :eng.D].permute(0, 2, 1).reshape(NFB, -1) \
            .contiguous()
        grads[name + ".bias"] = dbe
        if sink is not None:
            sink("head", grads)
        if tape["train_backbone"]:
            backward_backbone(eng, tape, dcrops, grads)
            if sink is not None:
                sink("backbone", grads)
    return grads
