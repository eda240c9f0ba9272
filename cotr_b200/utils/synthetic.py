"""Seeded synthetic weights, inputs and images (no checkpoint and no dataset are available offline).

`make_state_dict` builds a well-conditioned 381-entry `state_dict` for the COTR schema (SURVEY.md appendix C / E.3:
the reference initialisation gives predictions whose spread over queries is below the 1e-3 parity bar, so the q/k and
head gains are raised), `make_inputs` a normalised canvas + queries, `synthetic_image` a smooth uint8 test image.
Used by bench.py, the bring-up tools and - through `oracle/fixtures.py`, which re-exports these names - by the tests
and the golden-vector generators.  Everything is numpy `RandomState`-seeded, so the goldens stay reproducible.
"""
import numpy as np

D = 256
FF = 1024


def schema():
    """The 381-entry state_dict schema of the reference model as [(key, shape)]."""
    out = []
    body = "backbone.0.body"

    def bn(prefix, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{prefix}.{n}", (c,)))

    out.append((f"{body}.conv1.weight", (64, 3, 7, 7)))
    bn(f"{body}.bn1", 64)
    inplanes = 64
    for name, n_blocks, planes in (("layer1", 3, 64), ("layer2", 4, 128), ("layer3", 6, 256)):
        for i in range(n_blocks):
            p = f"{body}.{name}.{i}"
            out.append((f"{p}.conv1.weight", (planes, inplanes, 1, 1)))
            bn(f"{p}.bn1", planes)
            out.append((f"{p}.conv2.weight", (planes, planes, 3, 3)))
            bn(f"{p}.bn2", planes)
            out.append((f"{p}.conv3.weight", (planes * 4, planes, 1, 1)))
            bn(f"{p}.bn3", planes * 4)
            if i == 0:
                out.append((f"{p}.downsample.0.weight", (planes * 4, inplanes, 1, 1)))
                bn(f"{p}.downsample.1", planes * 4)
            inplanes = planes * 4
    out.append(("input_proj.weight", (D, 1024, 1, 1)))
    out.append(("input_proj.bias", (D,)))

    def mha(prefix):
        out.append((f"{prefix}.in_proj_weight", (3 * D, D)))
        out.append((f"{prefix}.in_proj_bias", (3 * D,)))
        out.append((f"{prefix}.out_proj.weight", (D, D)))
        out.append((f"{prefix}.out_proj.bias", (D,)))

    def ffn_ln(prefix, norms):
        out.append((f"{prefix}.linear1.weight", (FF, D)))
        out.append((f"{prefix}.linear1.bias", (FF,)))
        out.append((f"{prefix}.linear2.weight", (D, FF)))
        out.append((f"{prefix}.linear2.bias", (D,)))
        for n in norms:
            out.append((f"{prefix}.{n}.weight", (D,)))
            out.append((f"{prefix}.{n}.bias", (D,)))

    for l in range(6):
        p = f"transformer.encoder.layers.{l}"
        mha(f"{p}.self_attn")
        ffn_ln(p, ("norm1", "norm2"))
    for l in range(6):
        p = f"transformer.decoder.layers.{l}"
        mha(f"{p}.multihead_attn")
        ffn_ln(p, ("norm1", "norm2", "norm3"))
    out.append(("transformer.decoder.norm.weight", (D,)))
    out.append(("transformer.decoder.norm.bias", (D,)))
    for i, shp in enumerate(((D, D), (D, D), (2, D))):
        out.append((f"corr_embed.layers.{i}.weight", shp))
        out.append((f"corr_embed.layers.{i}.bias", (shp[0],)))
    return out


def make_state_dict(seed=0, qk_gain=3.0, head_gain=1.35, stem_gain=1.0):
    """numpy float32 state dict.  Conv: He-normal (fan_out); FrozenBN: random stats; transformer/head:
    xavier-uniform with q/k rows scaled by `qk_gain` and head matrices by `head_gain`; small biases.
    `stem_gain` multiplies the 7x7 stem kernel and divides `input_proj.weight`: the whole ResNet body then runs at
    ~stem_gain times the usual activation magnitude (ReLU networks are positively homogeneous up to the FrozenBN
    shifts) while the transformer sees ordinary values - the fixture that stresses the fp16 hi/lo storage range."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape in schema():
        if key.startswith("backbone") and len(shape) == 4:      # conv kernels
            fan_out = shape[0] * shape[2] * shape[3]
            w = rs.standard_normal(shape) * np.sqrt(2.0 / fan_out)
        elif key.startswith("backbone"):                       # FrozenBN buffers
            if key.endswith("running_var") or key.endswith(".weight"):
                w = rs.uniform(0.75, 1.25, shape)
            else:
                w = rs.standard_normal(shape) * 0.1
        elif key == "input_proj.weight":
            bound = np.sqrt(1.0 / shape[1])
            w = rs.uniform(-bound, bound, shape)
        elif len(shape) == 2:
            bound = np.sqrt(6.0 / (shape[0] + shape[1]))
            w = rs.uniform(-bound, bound, shape)
            if key.endswith("in_proj_weight"):
                w[: 2 * D] *= qk_gain
            if key.startswith("corr_embed"):
                w *= head_gain
        elif "norm" in key and key.endswith(".weight"):
            w = rs.uniform(0.8, 1.2, shape)
        else:                                                   # biases
            w = rs.standard_normal(shape) * 0.05
        sd[key] = np.ascontiguousarray(w, dtype=np.float32)
    if stem_gain != 1.0:
        sd["backbone.0.body.conv1.weight"] = np.ascontiguousarray(sd["backbone.0.body.conv1.weight"] * np.float32(stem_gain))
        sd["input_proj.weight"] = np.ascontiguousarray(sd["input_proj.weight"] / np.float32(stem_gain))
    return sd


def make_inputs(seed=1, batch=1, n_queries=1024):
    """img ~ N(0,1) (B,3,256,512) fp32, queries ~ U[0,1)^2 (B,Q,2) fp32."""
    rs = np.random.RandomState(seed)
    img = rs.standard_normal((batch, 3, 256, 512)).astype(np.float32)
    queries = rs.uniform(0.0, 1.0, (batch, n_queries, 2)).astype(np.float32)
    return img, queries


def synthetic_image(seed, h, w):
    """Smooth seeded uint8 RGB texture (low-frequency random field + gradient)."""
    rs = np.random.RandomState(seed)
    import cv2
    small = rs.uniform(0, 255, (h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
    img += np.linspace(-30, 30, w, dtype=np.float32)[None, :, None]
    return np.clip(img, 0, 255).astype(np.uint8)
