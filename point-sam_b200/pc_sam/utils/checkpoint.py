"""Checkpoint loading for the reference's ``model.safetensors`` files (evaluation/inference.py:11,48 uses
``safetensors.torch.load_model``).  The container format is small enough to read directly: an 8-byte little-endian
header length, a JSON table ``name -> {dtype, shape, data_offsets}`` and one flat byte buffer.  This reader has no
dependency on the ``safetensors`` package and maps the file instead of copying it twice."""
from __future__ import annotations

import json
import mmap
import struct
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

_DTYPES = {"F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16, "I64": torch.int64,
           "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}
_NAMES = {v: k for k, v in _DTYPES.items()}


def load_file(path: str) -> Dict[str, torch.Tensor]:
    out = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        table = json.loads(f.read(hlen).decode("utf-8"))
        base = 8 + hlen
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    for name, meta in table.items():
        if name == "__metadata__":
            continue
        if meta["dtype"] not in _DTYPES:
            raise ValueError(f"unsupported safetensors dtype {meta['dtype']} for {name}")
        b, e = meta["data_offsets"]
        dt = _DTYPES[meta["dtype"]]
        if e == b:
            out[name] = torch.empty(meta["shape"], dtype=dt)
            continue
        raw = np.frombuffer(mm, dtype=np.uint8, count=e - b, offset=base + b)
        out[name] = torch.from_numpy(raw.copy()).view(dt).reshape(meta["shape"])
    return out


def save_file(tensors: Dict[str, torch.Tensor], path: str, metadata=None):
    table, off, blobs = {}, 0, []
    if metadata:
        table["__metadata__"] = metadata
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        table[name] = {"dtype": _NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    head = json.dumps(table, separators=(",", ":")).encode("utf-8")
    head += b" " * ((8 - len(head) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(head)))
        f.write(head)
        for b in blobs:
            f.write(b)


def load_model(model: torch.nn.Module, path: str, strict: bool = True) -> Tuple[Iterable[str], Iterable[str]]:
    """Same contract as safetensors.torch.load_model: strict key match (RuntimeError listing the differences), returns
    (missing, unexpected).  Keys are the reference's own (timm ``transformer.blocks.N...``, apex FusedLayerNorm
    ``weight``/``bias``), which the mirror modules reproduce."""
    state = load_file(path)
    res = model.load_state_dict(state, strict=False)
    missing, unexpected = list(res.missing_keys), list(res.unexpected_keys)
    if strict and (missing or unexpected):
        raise RuntimeError(f"Error(s) in loading state_dict: missing {sorted(missing)[:8]}{'...' if len(missing) > 8 else ''}, "
                           f"unexpected {sorted(unexpected)[:8]}{'...' if len(unexpected) > 8 else ''}")
    return missing, unexpected
