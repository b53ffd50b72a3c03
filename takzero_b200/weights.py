"""Random-init weights of the reference architecture (takzero/src/network/net6_simhash.rs:43-141,
net4_simhash.rs, net5.rs) as the named f32 tensors tz_set_weights takes.

torch is used only as the initialiser: nn.Conv2d / nn.Linear default to the same Kaiming-uniform
scheme tch's `nn::conv2d` / `nn::linear` use, BatchNorm starts at weight 1 / bias 0 / mean 0 / var 1
(BN statistics of a freshly initialised network)."""
from __future__ import annotations

from typing import Dict

import numpy as np

FILTERS = 256


def res_blocks_for(n: int) -> int:
    return 20 if n == 5 else 16  # net5.rs:45 vs net4/net6 CORE_RES_BLOCKS


def channels(n: int):
    return 2 * (2 * n + 3 + 2) + 2, 3 + 4 * ((1 << n) - 2)  # repr.rs:103-108,133-139


def random_init(n: int, seed: int = 123, blocks: int | None = None) -> Dict[str, np.ndarray]:
    import torch
    from torch import nn

    blocks = res_blocks_for(n) if blocks is None else blocks
    cin, cout = channels(n)
    torch.manual_seed(seed)
    out: Dict[str, np.ndarray] = {}

    def conv(name, ci, co, k, bias):
        m = nn.Conv2d(ci, co, k, padding=k // 2, bias=bias)
        out[f"{name}.weight"] = m.weight.detach().numpy().copy()
        if bias:
            out[f"{name}.bias"] = m.bias.detach().numpy().copy()

    def bn(name):
        out[f"{name}.weight"] = np.ones(FILTERS, np.float32)
        out[f"{name}.bias"] = np.zeros(FILTERS, np.float32)
        out[f"{name}.running_mean"] = np.zeros(FILTERS, np.float32)
        out[f"{name}.running_var"] = np.ones(FILTERS, np.float32)

    conv("core.input_conv2d", cin, FILTERS, 3, False)
    bn("core.batch_norm")
    for b in range(blocks):
        for j in range(2):
            conv(f"core.res_block_{b}.{j}.conv2d", FILTERS, FILTERS, 3, False)
            bn(f"core.res_block_{b}.{j}.batch_norm")
    conv("policy.conv2d", FILTERS, cout, 3, True)
    for head in ("value", "ube"):
        conv(f"{head}.conv2d", FILTERS, 1, 1, True)
        lin = nn.Linear(n * n, 1)
        out[f"{head}.linear.weight"] = lin.weight.detach().numpy().copy()
        out[f"{head}.linear.bias"] = lin.bias.detach().numpy().copy()
    return out


def flops_per_position(n: int, blocks: int | None = None) -> float:
    """Algorithmic FLOPs of one network evaluation."""
    blocks = res_blocks_for(n) if blocks is None else blocks
    cin, cout = channels(n)
    nn_ = n * n
    mac = nn_ * 9 * cin * FILTERS + blocks * 2 * nn_ * 9 * FILTERS * FILTERS + nn_ * 9 * FILTERS * cout
    mac += 2 * (nn_ * FILTERS + nn_)
    return 2.0 * mac


def save_tzw(path: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write the TZW1 container the C++ hosts read (include/takzero_b200.hpp `Weights::load`)."""
    import struct

    with open(path, "wb") as f:
        f.write(b"TZW1" + struct.pack("<I", len(tensors)))
        for name, t in tensors.items():
            a = np.ascontiguousarray(t, dtype=np.float32)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<I", a.ndim))
            f.write(struct.pack(f"<{a.ndim}q", *a.shape))
            f.write(a.tobytes())


def tch_names(tensors: Dict[str, np.ndarray], first_suffix: int = 0) -> Dict[str, np.ndarray]:
    """Rename this package's tensor names to the ones tch's `VarStore` gives the reference network: both
    `SmallBlock`s of a `ResidualBlock` are built on one path (network/residual.rs:52-54), so half 0 keeps
    `core.res_block_B.conv2d.weight` and half 1 gets the collision suffix `__K` (K = variables in the store when the
    clash happened; the reader ignores the value, `first_suffix` only makes test files look real)."""
    import re

    out: Dict[str, np.ndarray] = {}
    k = first_suffix
    for name, t in tensors.items():
        m = re.fullmatch(r"(core\.res_block_\d+)\.([01])\.(.+)", name)
        if m:
            name = f"{m.group(1)}.{m.group(3)}" + (f"__{k}" if m.group(2) == "1" else "")
        out[name] = t
        k += 1
    return out


def save_ot(path: str, tensors: Dict[str, np.ndarray], rename: bool = True) -> None:
    """Write a libtorch module archive with one tensor attribute per name -- the container tch's
    `VarStore::save` produces for `model_latest.ot` (ZIP: `<stem>/data.pkl` + `<stem>/data/<key>`).  Used to make
    test inputs for the library's `.ot` reader; torch is the writer, so this is tooling, not the product path."""
    import torch

    m = torch.jit.ScriptModule()
    for name, t in (tch_names(tensors) if rename else tensors).items():
        m._c._register_attribute(name, torch._C.TensorType.get(), torch.from_numpy(np.ascontiguousarray(t)))
    m._c.save(str(path))
