"""CPU restatement of ``COTR.forward`` (TEST ORACLE - not product code).

A pure function of a flat ``{state_dict key: tensor}`` mapping, written with
torch CPU primitives in fp32 or fp64.  Every stage cites the reference lines it
restates (paths relative to /root/reference):

  * COTR/models/cotr_model.py:26-40          top-level forward
  * COTR/models/backbone.py:46-56            FrozenBatchNorm2d
  * COTR/models/backbone.py:79-92            left/right halves, cat on W
  * torchvision/models/resnet.py Bottleneck  (v1.5: stride on the 3x3)
  * COTR/models/position_encoding.py:41-45   lin_sine encoding
  * COTR/models/position_encoding.py:60-72   grid position embedding
  * COTR/models/transformer.py:47-58         flatten, tgt = 0
  * COTR/models/transformer.py:143-159       encoder layer (post-LN)
  * COTR/models/transformer.py:185-201       decoder layer (cross-attn only)
  * COTR/models/transformer.py:90-119        final decoder LayerNorm
  * COTR/models/position_encoding.py:14-26   MLP head
  * torch/nn/functional.py multi_head_attention_forward (packed in_proj,
    q scaled by head_dim**-0.5, softmax over keys, out_proj)

Pinned against the real reference by oracle/make_golden.py (max |delta| in fp64
recorded in tests/golden/*.npz and asserted by tests/test_oracle.py).
"""
import math

import torch
import torch.nn.functional as F

MAX_SIZE = 256            # COTR/utils/constants.py:2
D_MODEL = 256
N_HEAD = 8
HEAD_DIM = D_MODEL // N_HEAD
N_ENC = 6
N_DEC = 6
GRID_H, GRID_W = 16, 32
BN_EPS = 1e-5             # backbone.py:52
LN_EPS = 1e-5             # nn.LayerNorm default
RESNET_BLOCKS = (("layer1", 3, 64, 1), ("layer2", 4, 128, 2), ("layer3", 6, 256, 2))


def _frozen_bn(x, sd, prefix):
    # backbone.py:46-56: y = x * (w * rsqrt(rv + eps)) + (b - rm * w * rsqrt(rv + eps))
    w = sd[prefix + ".weight"]
    b = sd[prefix + ".bias"]
    rv = sd[prefix + ".running_var"]
    rm = sd[prefix + ".running_mean"]
    scale = w * (rv + BN_EPS).rsqrt()
    shift = b - rm * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def _bottleneck(x, sd, prefix, stride, has_downsample):
    # torchvision Bottleneck: 1x1 -> bn -> relu -> 3x3(stride) -> bn -> relu -> 1x1 -> bn -> (+id) -> relu
    y = F.conv2d(x, sd[prefix + ".conv1.weight"])
    y = F.relu(_frozen_bn(y, sd, prefix + ".bn1"))
    y = F.conv2d(y, sd[prefix + ".conv2.weight"], stride=stride, padding=1)
    y = F.relu(_frozen_bn(y, sd, prefix + ".bn2"))
    y = F.conv2d(y, sd[prefix + ".conv3.weight"])
    y = _frozen_bn(y, sd, prefix + ".bn3")
    if has_downsample:
        idn = F.conv2d(x, sd[prefix + ".downsample.0.weight"], stride=stride)
        idn = _frozen_bn(idn, sd, prefix + ".downsample.1")
    else:
        idn = x
    return F.relu(y + idn)


def backbone_half(x, sd, prefix="backbone.0.body"):
    """ResNet-50 up to layer3 on one (N,3,256,256) half -> (N,1024,16,16)."""
    y = F.conv2d(x, sd[prefix + ".conv1.weight"], stride=2, padding=3)
    y = F.relu(_frozen_bn(y, sd, prefix + ".bn1"))
    y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
    for name, n_blocks, _planes, stride in RESNET_BLOCKS:
        for i in range(n_blocks):
            y = _bottleneck(y, sd, f"{prefix}.{name}.{i}", stride if i == 0 else 1, i == 0)
    return y


def lin_sine(p, depth):
    """position_encoding.py:41-45 with bases 1..depth: cat([sin(k*pi*p)]_k + [cos(k*pi*p)]_k, -1)."""
    parts = [torch.sin(k * math.pi * p) for k in range(1, depth + 1)]
    parts += [torch.cos(k * math.pi * p) for k in range(1, depth + 1)]
    return torch.cat(parts, dim=-1)


def grid_position(dtype):
    """position_encoding.py:60-72 for the all-False mask of a 16x32 grid -> (512, 256), token n = i*32+j."""
    ones = torch.ones(1, GRID_H, GRID_W, dtype=torch.float32)
    y_embed = ones.cumsum(1)
    x_embed = ones.cumsum(2)
    eps = 1e-6
    y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps)
    x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps)
    pos = torch.stack([x_embed, y_embed], dim=-1)              # (1,16,32,2) fp32 like the reference
    enc = lin_sine(pos, D_MODEL // 4)                           # (1,16,32,256)
    return enc.reshape(GRID_H * GRID_W, D_MODEL).to(dtype)      # backbone.py:121 casts to feature dtype


def _mha(q_in, k_in, v_in, sd, prefix):
    """q_in (B,Lq,256), k_in/v_in (B,Lk,256); packed in_proj rows [0:256]=q,[256:512]=k,[512:768]=v."""
    w = sd[prefix + ".in_proj_weight"]
    b = sd[prefix + ".in_proj_bias"]
    d = D_MODEL
    q = F.linear(q_in, w[0:d], b[0:d]) * (HEAD_DIM ** -0.5)
    k = F.linear(k_in, w[d:2 * d], b[d:2 * d])
    v = F.linear(v_in, w[2 * d:3 * d], b[2 * d:3 * d])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, N_HEAD, HEAD_DIM).transpose(1, 2)
    k = k.view(B, Lk, N_HEAD, HEAD_DIM).transpose(1, 2)
    v = v.view(B, Lk, N_HEAD, HEAD_DIM).transpose(1, 2)
    attn = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _ln(x, sd, prefix):
    return F.layer_norm(x, (D_MODEL,), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def _ffn(x, sd, prefix):
    h = F.relu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]))
    return F.linear(h, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])


def encoder(src, pos, sd):
    """transformer.py:143-159 x6. src (B,512,256), pos (512,256)."""
    x = src
    for l in range(N_ENC):
        p = f"transformer.encoder.layers.{l}"
        qk = x + pos
        x = _ln(x + _mha(qk, qk, x, sd, p + ".self_attn"), sd, p + ".norm1")
        x = _ln(x + _ffn(x, sd, p), sd, p + ".norm2")
    return x


def decoder(memory, pos, qpos, sd):
    """transformer.py:185-201 x6 (no self-attention; norm1 unused) + decoder.norm. qpos (B,Q,256)."""
    t = torch.zeros_like(qpos)
    kmem = memory + pos
    for l in range(N_DEC):
        p = f"transformer.decoder.layers.{l}"
        t = _ln(t + _mha(t + qpos, kmem, memory, sd, p + ".multihead_attn"), sd, p + ".norm2")
        t = _ln(t + _ffn(t, sd, p), sd, p + ".norm3")
    return _ln(t, sd, "transformer.decoder.norm")


def head(hs, sd):
    """position_encoding.py:14-26 (3 layers, ReLU between) applied to the last decoder level only."""
    x = F.relu(F.linear(hs, sd["corr_embed.layers.0.weight"], sd["corr_embed.layers.0.bias"]))
    x = F.relu(F.linear(x, sd["corr_embed.layers.1.weight"], sd["corr_embed.layers.1.bias"]))
    return F.linear(x, sd["corr_embed.layers.2.weight"], sd["corr_embed.layers.2.bias"])


def cast_state_dict(sd, dtype):
    return {k: torch.as_tensor(v).to(dtype) for k, v in sd.items()}


@torch.no_grad()
def forward(sd, img, queries, dtype=torch.float32, return_intermediates=False):
    """sd: state_dict (any float dtype); img (B,3,256,512); queries (B,Q,2) -> pred (B,Q,2) in `dtype`."""
    sd = cast_state_dict(sd, dtype)
    img = torch.as_tensor(img).to(dtype)
    q_in = torch.as_tensor(queries).to(dtype)
    assert img.shape[-2:] == (MAX_SIZE, 2 * MAX_SIZE)            # backbone.py:80
    B = img.shape[0]
    left = backbone_half(img[..., :MAX_SIZE], sd)
    right = backbone_half(img[..., MAX_SIZE:], sd)
    feat = torch.cat([left, right], dim=-1)                      # (B,1024,16,32)  backbone.py:85
    src = F.conv2d(feat, sd["input_proj.weight"], sd["input_proj.bias"])
    src = src.flatten(2).transpose(1, 2)                         # (B,512,256), n = i*32+j
    pos = grid_position(dtype)
    qpos = lin_sine(q_in.reshape(-1, 2), D_MODEL // 4).reshape(B, -1, D_MODEL)
    mem = encoder(src, pos, sd)
    hs = decoder(mem, pos, qpos, sd)
    pred = head(hs, sd)
    if return_intermediates:
        return pred, {"feat": feat, "src": src, "pos": pos, "qpos": qpos, "mem": mem, "hs": hs}
    return pred
