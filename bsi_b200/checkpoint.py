"""Ingest checkpoints written by the reference's Lightning harness.

A reference `.ckpt` holds `state_dict` with the online denoiser under `model.*` and the EMA copy under
`ema_model.ema_model.*` (bsi/tasks/bsi.py:105-118), and the resolved run configuration under `config`
(bsi/lightning/callbacks.py:15-16).  `from_reference_checkpoint` rebuilds the native denoiser + BSI from that dict;
the fp32 weights are packed to the bf16 arena lazily on the first call, like any other `load_state_dict`.
"""

from __future__ import annotations

from typing import Any, Mapping

import torch

from .bsi import BSI, Discretization
from .models import DenoisingDiT, DenoisingVDMUNet, NyquistPositionalEmbedding
from .nn import FourierFeatures

_PREFIX = {"online": "model.", "ema": "ema_model.ema_model."}


def denoiser_state_dict(ckpt: Mapping[str, Any], which: str = "ema") -> dict[str, torch.Tensor]:
    """Denoiser weights of a reference checkpoint with the prefix stripped; `which` is "ema" (what eval uses) or "online"."""
    prefix = _PREFIX[which]
    sd = {k[len(prefix):]: v for k, v in ckpt["state_dict"].items() if k.startswith(prefix)}
    if not sd and which == "ema":  # runs without EMA evaluate the online weights (bsi/tasks/bsi.py:121-128)
        return denoiser_state_dict(ckpt, "online")
    if not sd:
        raise KeyError(f"checkpoint has no '{prefix}*' entries")
    return sd


def _data_shape(config: Mapping[str, Any]) -> tuple[int, int, int]:
    """``datamodule.data_shape()`` of the run (bsi/tasks/bsi.py:103) from the resolved data config, with the data modules'
    own rules: ImageNetDataModule -> (3, n, n) (bsi/data/imagenet.py:148-149, config/data/imagenet.yaml ``n``);
    CIFAR10DataModule -> (3, 32, 32) (bsi/data/cifar10.py:148-149).  The weights cannot tell: the DiT's positional table is
    a non-persistent buffer, so the token grid is not in the state_dict."""
    data = config["data"]
    target = str(data.get("_target_", ""))
    if target.endswith("ImageNetDataModule") or ("n" in data and not target):
        n = int(data["n"])
        return (3, n, n)
    if target.endswith("CIFAR10DataModule"):
        return (3, int(data.get("height", 32)), int(data.get("width", 32)))
    if "data_shape" in data:  # explicit override for checkpoints of other data modules
        return tuple(int(v) for v in data["data_shape"])
    raise KeyError(f"cannot derive the data shape from the checkpoint's data config (_target_={target!r}); add data.data_shape")


def build_denoiser(model_cfg: Mapping[str, Any], data_shape) -> torch.nn.Module:
    """Instantiate the native counterpart of a reference `task.model` config node (config/task/model/bsi/{dit,unet}.yaml)."""
    cfg = dict(model_cfg)
    target = cfg.pop("_target_")
    ff_cfg = cfg.pop("fourier_features", None)
    ff = FourierFeatures(n_min=ff_cfg["n_min"], n_max=ff_cfg["n_max"]) if ff_cfg else None
    if target.endswith("DenoisingDiT"):
        return DenoisingDiT(data_shape, cfg["patch_size"], cfg["dim"], cfg["depth"], cfg["heads"], dropout=cfg.get("dropout"), fourier_features=ff)
    if target.endswith("DenoisingVDMUNet"):
        pe = cfg.pop("pos_emb")
        pos = NyquistPositionalEmbedding.from_config(pe["size"], pe["expected_rate"])
        return DenoisingVDMUNet(data_shape, pos, cfg["actfn"], cfg["dim"], cfg["levels"], cfg["pos_emb_mult"], n_attention_heads=cfg.get("n_attention_heads", 1),
                                dropout=cfg.get("dropout"), downsampling_attention=cfg.get("downsampling_attention", False), fourier_features=ff,
                                padding_mode=cfg.get("padding_mode", "zeros"))
    raise NotImplementedError(f"no native implementation of {target}")


def from_reference_checkpoint(ckpt: Mapping[str, Any] | str, which: str = "ema", device="cuda"):
    """(BSI, denoiser) rebuilt from a reference checkpoint dict or path, ready for sample / elbo on `device`."""
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    config = ckpt["config"]
    shape = _data_shape(config)
    model = build_denoiser(config["task"]["model"], shape)
    model.load_state_dict(denoiser_state_dict(ckpt, which))
    model = model.to(device).eval().requires_grad_(False)
    b = {k: v for k, v in dict(config["task"]["bsi"]).items() if k != "_target_"}
    bsi = BSI(model, data_shape=shape, discretization=Discretization.image_8bit(), **b).to(device)
    return bsi, model
