"""Deterministic, platform-independent synthetic weights / data for parity tests.

Values come from a numpy Philox4x32 stream keyed by a CRC of the parameter name, so the
golden generator (build container, real reference) and the GPU tests (fresh box, no
reference) rebuild bit-identical tensors without shipping them.
"""

from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bsi_oracle as O  # noqa: E402  (tests may import the oracle)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REFERENCE_DIR = "/root/reference"


def det_uniform(tag: str, shape, seed: int = 0) -> torch.Tensor:
    """U(-1,1) fp32 tensor, a pure function of (tag, shape, seed)."""
    n = int(np.prod(shape)) if len(shape) else 1
    quads = (n + 3) // 4
    ctr = np.zeros((quads, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(quads, dtype=np.uint32)
    key = np.array([zlib.crc32(tag.encode()) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)
    r = O.philox4x32(ctr, np.broadcast_to(key, (quads, 2))).reshape(-1)[:n]
    u = (r.astype(np.float64) + 0.5) * 2.0**-32
    return torch.from_numpy((2 * u - 1).astype(np.float32)).reshape(tuple(shape))


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def det_state_dict(shapes: dict[str, tuple], seed: int = 0, bf16_exact: bool = True) -> dict[str, torch.Tensor]:
    """Deterministic non-degenerate weights for a {name: shape} map.

    matrices / conv kernels: U(-1,1)/sqrt(fan_in);  1-d '*.weight' (norm gains): 1 + 0.1 U;
    biases: 0.05 U.  adaLN-Zero output layers are therefore NOT zero (a
    default-initialised DiT is an identity stack).
    """
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        u = det_uniform(name, shape, seed)
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            w = u / np.sqrt(fan_in)
        elif name.endswith("weight"):
            w = 1 + 0.1 * u
        else:
            w = 0.05 * u
        out[name] = bf16_round(w) if bf16_exact else w
    return out


def det_images(tag: str, batch: int, shape, seed: int = 0) -> torch.Tensor:
    """8-bit grid images as the reference's data modules produce them: u8 * (2/255) - 1 (bsi/data/imagenet.py:56)."""
    u = det_uniform(tag, (batch, *shape), seed)
    u8 = ((u + 1) * 128).floor().clamp(0, 255)
    return u8.to(torch.float32) * (2 / 255) - 1


def dit_shapes(spec: "O.DiTSpec") -> dict[str, tuple]:
    p2 = spec.patch**2
    d = spec.dim
    shapes = {
        "dit.patch_encoder.weight": (d, p2 * spec.in_channels),
        "dit.patch_encoder.bias": (d,),
        "dit.patch_decoder.0.weight": (d,),
        "dit.patch_decoder.0.bias": (d,),
        "dit.patch_decoder.1.weight": (p2 * spec.data_shape[0], d),
        "dit.patch_decoder.1.bias": (p2 * spec.data_shape[0],),
    }
    for i in range(spec.depth):
        b = f"dit.blocks.{i}."
        shapes.update(
            {
                b + "attn.to_qkv.weight": (3 * d, d),
                b + "attn.to_qkv.bias": (3 * d,),
                b + "attn.to_out.weight": (d, d),
                b + "attn.to_out.bias": (d,),
                b + "mlp.0.weight": (4 * d, d),
                b + "mlp.0.bias": (4 * d,),
                b + "mlp.2.weight": (d, 4 * d),
                b + "mlp.2.bias": (d,),
                b + "adaLN_modulation.0.weight": (d, d),
                b + "adaLN_modulation.0.bias": (d,),
                b + "adaLN_modulation.2.weight": (6 * d, d),
                b + "adaLN_modulation.2.bias": (6 * d,),
            }
        )
    return shapes


def unet_shapes(spec: "O.UNetSpec", dropout: bool = True) -> dict[str, tuple]:
    d = spec.dim
    cdim = spec.pos_size * spec.pos_mult
    cin = spec.data_shape[0] * (1 if spec.fourier is None else 1 + 2 * (spec.fourier[1] - spec.fourier[0] + 1))
    second = "layers.6" if dropout else "layers.5"
    shapes = {
        "pos_map.1.weight": (cdim, spec.pos_size),
        "pos_map.1.bias": (cdim,),
        "pos_map.3.weight": (cdim, cdim),
        "pos_map.3.bias": (cdim,),
        "encode.weight": (d, cin, 3, 3),
        "encode.bias": (d,),
        "decode.weight": (spec.data_shape[0], d, 1, 1),
        "decode.bias": (spec.data_shape[0],),
        "u_net.center_block.1.fn.0.weight": (d,),
        "u_net.center_block.1.fn.0.bias": (d,),
        "u_net.center_block.1.fn.1.to_qkv.weight": (3 * d, d, 3, 3),
        "u_net.center_block.1.fn.1.to_qkv.bias": (3 * d,),
        "u_net.center_block.1.fn.1.to_out.weight": (d, d, 3, 3),
        "u_net.center_block.1.fn.1.to_out.bias": (d,),
    }

    def res(pre, cin_):
        shapes.update(
            {
                pre + ".project_onto_scale_shift.weight": (2 * d, cdim),
                pre + ".project_onto_scale_shift.bias": (2 * d,),
                pre + ".layers.0.weight": (cin_,),
                pre + ".layers.0.bias": (cin_,),
                pre + ".layers.2.weight": (d, cin_, 3, 3),
                pre + ".layers.2.bias": (d,),
                pre + f".{second}.weight": (d, d, 3, 3),
                pre + f".{second}.bias": (d,),
            }
        )
        if cin_ != d:
            shapes.update({pre + ".skip.weight": (d, cin_, 1, 1), pre + ".skip.bias": (d,)})

    for i in range(spec.levels):
        res(f"u_net.downsampling_blocks.{i}.0", d)
        res(f"u_net.upsampling_blocks.{i}.0", 2 * d)
    res("u_net.center_block.0", d)
    res("u_net.center_block.2", d)
    return shapes


TOY_SHAPES = {"layer.weight": (3, 4, 3, 3), "layer.bias": (3,)}


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name), map_location="cpu", weights_only=True)


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_DIR, "bsi"))


# optimizer-side fixture (tests/golden/optim.pt): odd sizes so the flat arena needs padding, gradients scaled so that the
# global-norm clip is active on some steps only, EMA schedule short enough to pass through init / copy / lerp / skipped steps
OPTIM_SHAPES = [(7, 5), (33,), (4, 4, 3), (1,)]
OPTIM_HYPER = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
OPTIM_MAX_NORM = 1.0
OPTIM_EMA = dict(beta=0.9999, update_after_step=4, update_every=2)
OPTIM_STEPS = 14


def optim_grad(i: int, step: int) -> torch.Tensor:
    scale = (0.02, 0.3, 1.5)[step % 3]
    return det_uniform(f"optim.g{i}.{step}", OPTIM_SHAPES[i]) * scale


def _mix32(x: torch.Tensor) -> torch.Tensor:
    """csrc/common.cuh mix32 on int64 tensors holding uint32 values."""
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & m
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & m
    return x ^ (x >> 16)


def dropout_keep(seed: int, idx: torch.Tensor, p: float) -> torch.Tensor:
    """Keep mask of the native kernels' stateless dropout (csrc/common.cuh dropout_keep) for element indices `idx` (int64)."""
    thresh = int(float(np.float32(p)) * 65536.0)
    h = _mix32(((idx >> 1) * 0x9E3779B9 + seed) & 0xFFFFFFFF)  # 16 bits per element: elements 2i, 2i + 1 share the hash of pair i
    return torch.where((idx & 1) == 1, h >> 16, h & 0xFFFF) >= thresh


def attention_dropout_mask(seed: int, B: int, heads: int, T: int, p: float) -> torch.Tensor:
    """[B, heads, T(query), T(key)] keep mask of bsi_attention_dropout_bf16 / bsi_attention_backward_bf16."""
    bh = torch.arange(B * heads, dtype=torch.int64)
    sd = _mix32((seed ^ ((bh * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF)
    idx = torch.arange(T * T, dtype=torch.int64)
    thresh = int(float(np.float32(p)) * 65536.0)
    h = _mix32(((idx[None, :] >> 1) * 0x9E3779B9 + sd[:, None]) & 0xFFFFFFFF)
    keep = torch.where((idx[None, :] & 1) == 1, h >> 16, h & 0xFFFF) >= thresh
    return keep.reshape(B, heads, T, T)


def import_reference():
    """Import the real reference package (build container only)."""
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    sys.dont_write_bytecode = True
    import bsi.bsi as ref_bsi  # noqa
    import bsi.models.dit as ref_dit  # noqa
    import bsi.models.vdm_unet as ref_unet  # noqa
    import bsi.models.pos_emb as ref_pos  # noqa
    import bsi.nn as ref_nn  # noqa

    return ref_bsi, ref_dit, ref_unet, ref_pos, ref_nn
