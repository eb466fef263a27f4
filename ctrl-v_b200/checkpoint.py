"""diffusers-format checkpoint ingest.

The reference loads its networks with `ControlNetModel.from_pretrained(path, subfolder="controlnet")`
and `UNetSpatioTemporalConditionModel.from_pretrained(path, subfolder="unet")`
(/root/reference/tools/eval_video_controlnet.py:114-118, tools/eval_overall.py:203-214) and saves
them with accelerate / `save_pretrained` (tools/train_video_controlnet.py:151-182).  On disk that
is a directory per network:

    <path>/<subfolder>/config.json
    <path>/<subfolder>/diffusion_pytorch_model[.<variant>].safetensors        (or)
    <path>/<subfolder>/diffusion_pytorch_model[.<variant>].safetensors.index.json + shards  (or)
    <path>/<subfolder>/diffusion_pytorch_model[.<variant>].bin

This module reads and writes that layout without diffusers: a self-contained safetensors
reader/writer (8-byte little-endian header length, JSON header {name: {dtype, shape,
data_offsets}}, raw little-endian tensor bytes) and `torch.load(weights_only=True)` for `.bin`.
Host-side only: no CUDA, no kernels.
"""
from __future__ import annotations

import json
import os
import struct
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch

WEIGHTS_NAME = "diffusion_pytorch_model"
CONFIG_NAME = "config.json"

_ST_DTYPES = {
    "F64": (torch.float64, np.float64), "F32": (torch.float32, np.float32),
    "F16": (torch.float16, np.float16), "BF16": (torch.bfloat16, np.uint16),
    "I64": (torch.int64, np.int64), "I32": (torch.int32, np.int32), "I16": (torch.int16, np.int16),
    "I8": (torch.int8, np.int8), "U8": (torch.uint8, np.uint8), "BOOL": (torch.bool, np.bool_),
}
_ST_NAMES = {v[0]: k for k, v in _ST_DTYPES.items()}


def read_safetensors(path: str) -> "OrderedDict[str, torch.Tensor]":
    """Tensors of one .safetensors file (memory-mapped, copied out per tensor)."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than 8 bytes)")
        (n,) = struct.unpack("<Q", head)
        if n > size - 8:
            raise ValueError(f"{path}: header length {n} exceeds the file size {size}")
        meta = json.loads(f.read(n).decode("utf-8"))
    base = 8 + n
    mm = np.memmap(path, dtype=np.uint8, mode="r", offset=base) if size > base else np.zeros(0, np.uint8)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, info in meta.items():
        if name == "__metadata__":
            continue
        if info["dtype"] not in _ST_DTYPES:
            raise ValueError(f"{path}: tensor {name} has unsupported dtype {info['dtype']}")
        tdt, ndt = _ST_DTYPES[info["dtype"]]
        b0, b1 = info["data_offsets"]
        shape = tuple(info["shape"])
        count = int(np.prod(shape)) if shape else 1
        if b1 - b0 != count * np.dtype(ndt).itemsize or b1 > mm.size:
            raise ValueError(f"{path}: tensor {name} has inconsistent offsets {b0}:{b1} for shape {shape}")
        arr = np.array(mm[b0:b1]).view(ndt).reshape(shape) if count else np.zeros(shape, ndt)
        t = torch.from_numpy(arr)
        out[name] = t.view(torch.bfloat16) if tdt is torch.bfloat16 else t
    return out


def write_safetensors(path: str, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None):
    header = OrderedDict()
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    blobs = []
    off = 0
    for name in sorted(tensors):
        t = tensors[name].detach().to("cpu").contiguous()
        if t.dtype not in _ST_NAMES:
            raise ValueError(f"tensor {name}: dtype {t.dtype} cannot be stored as safetensors")
        raw = (t.view(torch.uint16) if t.dtype is torch.bfloat16 else t).numpy().tobytes()
        header[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def _weights_stem(variant: Optional[str]) -> str:
    return WEIGHTS_NAME + (f".{variant}" if variant else "")


def load_diffusers_dir(path: str, subfolder: Optional[str] = None, variant: Optional[str] = None
                       ) -> Tuple[dict, "OrderedDict[str, torch.Tensor]"]:
    """(config dict, state dict with diffusers key names) of one network directory."""
    d = os.path.join(path, subfolder) if subfolder else path
    if not os.path.isdir(d):
        raise OSError(f"{d} is not a directory (this build reads local checkpoints only; no hub download)")
    cfg_path = os.path.join(d, CONFIG_NAME)
    if not os.path.isfile(cfg_path):
        raise OSError(f"no {CONFIG_NAME} in {d}")
    with open(cfg_path, "r", encoding="utf-8") as f:
        config = json.load(f)
    stem = _weights_stem(variant)
    st, idx, binf = (os.path.join(d, stem + ".safetensors"), os.path.join(d, stem + ".safetensors.index.json"),
                     os.path.join(d, stem + ".bin"))
    if os.path.isfile(st):
        sd = read_safetensors(st)
    elif os.path.isfile(idx):
        with open(idx, "r", encoding="utf-8") as f:
            wm = json.load(f)["weight_map"]
        sd = OrderedDict()
        for shard in sorted(set(wm.values())):
            part = read_safetensors(os.path.join(d, shard))
            for k, v in part.items():
                if wm.get(k) == shard:
                    sd[k] = v
        missing = [k for k in wm if k not in sd]
        if missing:
            raise OSError(f"{idx}: shards do not contain {missing[:3]}")
    elif os.path.isfile(binf):
        sd = torch.load(binf, map_location="cpu", weights_only=True)
    else:
        raise OSError(f"no {stem}.safetensors / .safetensors.index.json / .bin in {d}")
    return config, sd


def save_diffusers_dir(path: str, config: dict, state_dict: Dict[str, torch.Tensor], class_name: str,
                       subfolder: Optional[str] = None, variant: Optional[str] = None,
                       safe_serialization: bool = True, max_shard_bytes: Optional[int] = None):
    d = os.path.join(path, subfolder) if subfolder else path
    os.makedirs(d, exist_ok=True)
    cfg = OrderedDict([("_class_name", class_name), ("_diffusers_version", "0.27.2")])
    for k, v in config.items():
        cfg[k] = list(v) if isinstance(v, tuple) else v
    with open(os.path.join(d, CONFIG_NAME), "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=2)
    stem = _weights_stem(variant)
    if not safe_serialization:
        torch.save(OrderedDict((k, v.detach().cpu()) for k, v in state_dict.items()), os.path.join(d, stem + ".bin"))
        return
    if not max_shard_bytes:
        write_safetensors(os.path.join(d, stem + ".safetensors"), state_dict, {"format": "pt"})
        return
    shards, cur, cur_b = [], OrderedDict(), 0
    for k in sorted(state_dict):
        nb = state_dict[k].numel() * state_dict[k].element_size()
        if cur and cur_b + nb > max_shard_bytes:
            shards.append(cur); cur, cur_b = OrderedDict(), 0
        cur[k] = state_dict[k]; cur_b += nb
    shards.append(cur)
    wm = OrderedDict()
    for i, sh in enumerate(shards):
        name = f"{stem}-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        write_safetensors(os.path.join(d, name), sh, {"format": "pt"})
        for k in sh:
            wm[k] = name
    with open(os.path.join(d, stem + ".safetensors.index.json"), "w", encoding="utf-8") as f:
        json.dump({"metadata": {}, "weight_map": wm}, f, indent=2)
