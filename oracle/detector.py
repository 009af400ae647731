"""Oracle: functional CPU restatement of the BALF detector forward.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Operates on a plain
``state_dict`` (name -> tensor) so that it shares no module code with the
product.  Parity: PINNED against ``/root/reference/balf/model`` by
``oracle/make_golden.py`` (fixtures in ``tests/golden/detector_*.npz``).

Reference followed:
  balf/model/mlp_ma_decoder.py:8-23    block / unblock rearranges
  balf/model/mlp_ma_decoder.py:25-70   grid gMLP layer + grid gating unit
  balf/model/mlp_ma_decoder.py:72-117  block gMLP layer + block gating unit
  balf/model/mlp_ma_decoder.py:119-149 residual split-head multi-axis gMLP
  balf/model/mlp_ma_decoder.py:151-199 squeeze-excite + residual channel attention
  balf/model/mlp_ma_decoder.py:201-244 Down
  balf/model/mlp_ma_decoder.py:246-285 MLP_MA_DECODER
  balf/model/decoder.py:5-30           DetectorHead
  balf/utils/tensor_op.py:1-27         pixel_shuffle
"""
import torch
import torch.nn.functional as F

P = "residual_split_head_multi_axis_gmlp_layer"
R = "residual_channel_attention_block"


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], 1e-5)


def to_tokens(x, fh, fw):
    """[n,h,w,c] -> [n, cells, within-cell, c]  (mlp_ma_decoder.py:8-16)."""
    n, h, w, c = x.shape
    gh, gw = h // fh, w // fw
    if gh * fh != h or gw * fw != w:
        raise ValueError("spatial size %dx%d not divisible by patch %dx%d" % (h, w, fh, fw))
    x = x.reshape(n, gh, fh, gw, fw, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, gh * gw, fh * fw, c)


def from_tokens(x, gh, gw, fh, fw):
    """inverse of to_tokens (mlp_ma_decoder.py:18-23)."""
    n, _, _, c = x.shape
    x = x.reshape(n, gh, gw, fh, fw, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, gh * fh, gw * fw, c)


def _gated_mlp(sd, pre, gate, x, mix_axis):
    """norm -> dense1 -> GELU -> gating unit -> dense2, residual (lines 57-70 / 104-117).

    x: [n, g, f, c].  mix_axis = 1 mixes across cells (grid), 2 mixes inside a cell (block).
    """
    y = F.gelu(_lin(sd, pre + ".dense1", _ln(sd, pre + ".norm", x)))
    c = y.shape[-1] // 2
    y1, y2 = y[..., :c], y[..., c:]
    y2 = _ln(sd, pre + "." + gate + ".norm", y2)
    w = sd[pre + "." + gate + ".dense.weight"]
    b = sd[pre + "." + gate + ".dense.bias"]
    if mix_axis == 1:      # lines 39-41: permute(0,3,2,1), Linear over the cell axis
        y2 = torch.einsum("pg,ngfc->npfc", w, y2) + b.view(1, -1, 1, 1)
    else:                  # lines 86-88: permute(0,1,3,2), Linear over the within-cell axis
        y2 = torch.einsum("qf,ngfc->ngqc", w, y2) + b.view(1, 1, -1, 1)
    y = _lin(sd, pre + ".dense2", y1 * (y2 + 1.0))
    return x + y


def down_forward(sd, pre, x, downsample, grid=(8, 8), block=(8, 8)):
    """One ``Down`` stage.  x: NCHW, returns NCHW (mlp_ma_decoder.py:223-244)."""
    x = x.permute(0, 2, 3, 1)
    x0 = F.relu(_lin(sd, pre + ".conv.0", x))
    n, h, w, c = x0.shape
    # multi-axis gMLP (lines 132-149)
    t = F.gelu(_lin(sd, pre + "." + P + ".dense1", _ln(sd, pre + "." + P + ".norm", x0)))
    u, v = t[..., :c], t[..., c:]
    fh, fw = h // grid[0], w // grid[1]
    u = from_tokens(_gated_mlp(sd, pre + "." + P + ".grid_gmlp_layer", "grid_gating_unit",
                               to_tokens(u, fh, fw), 1), grid[0], grid[1], fh, fw)
    gh, gw = h // block[0], w // block[1]
    v = from_tokens(_gated_mlp(sd, pre + "." + P + ".block_gmlp_layer", "block_gating_unit",
                               to_tokens(v, block[0], block[1]), 2), gh, gw, block[0], block[1])
    x1 = _lin(sd, pre + "." + P + ".dense2", torch.cat([u, v], -1)) + x0
    # residual channel attention (lines 185-199) with squeeze-excite (lines 164-171)
    r = _lin(sd, pre + "." + R + ".conv1", _ln(sd, pre + "." + R + ".norm", x1))
    r = _lin(sd, pre + "." + R + ".conv2", F.leaky_relu(r, 0.2))
    pooled = r.mean(dim=(1, 2))
    s = torch.sigmoid(_lin(sd, pre + "." + R + ".calayer.excite.2",
                           F.relu(_lin(sd, pre + "." + R + ".calayer.excite.0", pooled))))
    x2 = r * s.view(n, 1, 1, c) + x1
    out = (x2 + x0).permute(0, 3, 1, 2)
    if downsample:
        return F.max_pool2d(out, 2)
    return _lin(sd, pre + ".conv2", out.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def head_forward(sd, feat, cell=8, pre="detector_head"):
    """DetectorHead (decoder.py:16-30) + pixel_shuffle (tensor_op.py:1-27), eval-mode BN."""
    x = _lin(sd, pre + ".dense", F.relu(feat).permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    logits = F.batch_norm(x, sd[pre + ".norm.running_mean"], sd[pre + ".norm.running_var"],
                          sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], False, 0.0, 1e-5)
    p = torch.softmax(logits, dim=1)[:, :-1]
    n, _, hc, wc = p.shape
    prob = p.reshape(n, cell, cell, hc, wc).permute(0, 3, 1, 4, 2).reshape(n, hc * cell, wc * cell)
    return {"logits": logits, "prob": prob}


def detector_forward(sd, x, grid=(8, 8), block=(8, 8), cell=8):
    """MLP_MA_DECODER.forward (mlp_ma_decoder.py:278-285).  x: [B,Cin,H,W] fp32/fp64."""
    for i in (1, 2, 3):
        x = down_forward(sd, "down%d" % i, x, True, grid, block)
    x = down_forward(sd, "down4", x, False, grid, block)
    return head_forward(sd, x, cell)
