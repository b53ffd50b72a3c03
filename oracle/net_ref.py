"""CPU ORACLE (test infrastructure, NOT product code): f32 PyTorch restatement of the reference
network, used to check the CUDA network and as the forward pass of the restated CPU baseline.

Follows takzero/src/network/net6_simhash.rs:43-119,194-201,259-324 (net4_simhash.rs is the same
with N = 4; net5.rs:45 has 20 residual blocks) and residual.rs:13-63.  The reference builds these
layers through tch/libtorch, i.e. the same ATen kernels this file calls; the reference's tests pin
shapes only (net6_simhash.rs:340-367), so numerical parity of the network is "unpinned" and this
restatement is the de-facto oracle.  The SimHash set of a freshly initialised reference network is
empty, so `forward_hash` returns MAXIMUM_VARIANCE = 4.0 for every position (:243-256)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
from torch import nn

from . import oracle as O

FILTERS = 256
MAXIMUM_VARIANCE = 4.0


def res_blocks_for(n: int) -> int:
    return 20 if n == 5 else 16


class SmallBlock(nn.Module):  # residual.rs:13-43
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv2d = nn.Conv2d(cin, cout, 3, stride=1, padding=1, bias=False)
        self.batch_norm = nn.BatchNorm2d(cout, eps=1e-5, momentum=0.1)

    def forward(self, x):
        return self.batch_norm(self.conv2d(x))


class ResidualBlock(nn.Sequential):  # residual.rs:45-63
    def __init__(self, channels: int):
        super().__init__(SmallBlock(channels, channels), SmallBlock(channels, channels))

    def forward(self, x):
        return torch.relu(self[1](torch.relu(self[0](x))) + x)


class Core(nn.Module):  # net6_simhash.rs:43-72
    def __init__(self, cin: int, blocks: int):
        super().__init__()
        self.input_conv2d = nn.Conv2d(cin, FILTERS, 3, stride=1, padding=1, bias=False)
        self.batch_norm = nn.BatchNorm2d(FILTERS, eps=1e-5, momentum=0.1)
        for b in range(blocks):
            self.add_module(f"res_block_{b}", ResidualBlock(FILTERS))
        self.blocks = blocks

    def forward(self, x):
        x = torch.relu(self.batch_norm(self.input_conv2d(x)))
        for b in range(self.blocks):
            x = getattr(self, f"res_block_{b}")(x)
        return x


class Head(nn.Module):  # value_net / ube_net, net6_simhash.rs:88-119
    def __init__(self, n: int, tanh: bool):
        super().__init__()
        self.conv2d = nn.Conv2d(FILTERS, 1, 1, stride=1)
        self.linear = nn.Linear(n * n, 1)
        self.n, self.tanh = n, tanh

    def forward(self, x):
        y = self.linear(torch.relu(self.conv2d(x)).view(-1, self.n * self.n))
        return torch.tanh(y) if self.tanh else y


class Policy(nn.Module):  # net6_simhash.rs:74-86
    def __init__(self, cout: int):
        super().__init__()
        self.conv2d = nn.Conv2d(FILTERS, cout, 3, stride=1, padding=1)

    def forward(self, x):
        return self.conv2d(x)


class Rnd(nn.Module):  # net5.rs:120-146 `rnd`
    def __init__(self, size: int):
        super().__init__()
        self.input_linear = nn.Linear(size, 1024)
        self.hidden_linear = nn.Linear(1024, 1024)
        self.final_linear = nn.Linear(1024, 512)

    def forward(self, x):
        x = x.reshape(x.shape[0], -1)
        x = x / x.square().sum(dim=1, keepdim=True)
        return self.final_linear(torch.relu(self.hidden_linear(torch.relu(self.input_linear(x)))))


class Net(nn.Module):
    def __init__(self, n: int, seed: int = 123, blocks: int | None = None, randomize_bn: bool = False, rnd: bool = False):
        super().__init__()
        torch.manual_seed(seed)
        L = O.lib()
        self.n = n
        self.cin = L.tk_input_channels(n)
        self.cout = L.tk_output_channels(n)
        self.core = Core(self.cin, blocks if blocks is not None else res_blocks_for(n))
        self.policy = Policy(self.cout)
        self.value = Head(n, True)
        self.ube = Head(n, False)
        # root.randn_standard("simhash_matrix", [input_size, HASH_BITS]) (net6_simhash.rs:136-139)
        self.simhash_matrix = torch.randn(self.cin * n * n, 32, generator=torch.Generator().manual_seed(seed + 7))
        self.simhash_set: set = set()  # indices whose bit is set (the reference keeps a 2^32-bit BitBox)
        self.has_rnd = rnd
        if rnd:  # net5.rs:163-170: the 5x5 network's local uncertainty is RND, not a hash set
            self.rnd_learning = Rnd(self.cin * n * n)
            self.rnd_target = Rnd(self.cin * n * n)
            self.min = nn.Parameter(torch.zeros(1))
            self.max = nn.Parameter(torch.ones(1))
        if randomize_bn:  # exercise the BN folding with non-trivial statistics
            g = torch.Generator().manual_seed(seed + 1)
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
                    m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
                    m.running_mean.data = 0.2 * torch.randn(m.running_mean.shape, generator=g)
                    m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)
        self.eval()

    @torch.no_grad()
    def forward(self, xs):  # forward_t(xs, false), net6_simhash.rs:194-201
        core = self.core(xs)
        return self.policy(core), self.value(core), self.ube(core)

    @torch.no_grad()
    def simhash_dots(self, xs):
        """get_indices (net6_simhash.rs:203-234) up to the sign test: the side-to-move plane is zeroed."""
        xs = xs.clone()
        xs[:, self.cin - 2] = 0.0
        return xs.reshape(xs.shape[0], -1) @ self.simhash_matrix

    @torch.no_grad()
    def normalized_rnd(self, xs):
        """net5.rs:193-211: sum((learning - target)^2), normalized with min / max, times MAXIMUM_VARIANCE."""
        rnd = (self.rnd_learning(xs) - self.rnd_target(xs)).square().sum(dim=1)
        return torch.clamp((rnd - self.min) / (self.max - self.min), 0.0, 1.0) * MAXIMUM_VARIANCE

    def get_indices(self, xs) -> np.ndarray:
        dots = self.simhash_dots(xs).numpy()
        bits = (~(dots < 0.0)).astype(np.uint64)
        return (bits << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)

    @staticmethod
    @torch.no_grad()
    def lcghash_indices(xs: torch.Tensor, lcghash_init: torch.Tensor) -> np.ndarray:
        """`get_indices` of the LCG-hash network (net4_lcghash.rs:203-241), the same tensor operations:
        (xs * init) reinterpreted as i32, folded with acc = acc * MULTIPLIER + INCREMENT + x (wrapping i64) along
        the columns, then the rows, then the channels; index = |acc| >> (63 - HASH_BITS)."""
        MULTIPLIER, INCREMENT, HASH_BITS = 6_364_136_223_846_793_005, 1, 32
        batch, channels, rows, _cols = xs.shape
        parts = (xs * lcghash_init).view(torch.int32).split(1, 3)
        acc = torch.zeros(batch, channels, rows, dtype=torch.int64)
        for x in parts:
            acc *= MULTIPLIER
            acc += INCREMENT
            acc += x.squeeze(3)
        parts = acc.split(1, 2)
        acc = torch.zeros(batch, channels, dtype=torch.int64)
        for x in parts:
            acc *= MULTIPLIER
            acc += INCREMENT
            acc += x.squeeze(2)
        parts = acc.split(1, 1)
        acc = torch.zeros(batch, dtype=torch.int64)
        for x in parts:
            acc *= MULTIPLIER
            acc += INCREMENT
            acc += x.squeeze(1)
        return (acc.abs() >> (63 - HASH_BITS)).numpy().astype(np.uint32)

    def bitset_bytes(self) -> np.ndarray:
        """The reference's bitvec.bin image: 2^29 bytes, bit i of the set = byte i/8, bit i%8."""
        b = np.zeros(1 << 29, dtype=np.uint8)
        for i in self.simhash_set:
            b[i >> 3] |= 1 << (i & 7)
        return b

    def tensors(self) -> Dict[str, np.ndarray]:
        """Named f32 tensors in the layout tz_set_weights documents."""
        return {k: v.detach().cpu().numpy().astype(np.float32).copy() for k, v in self.state_dict().items()
                if not k.endswith("num_batches_tracked")}

    @torch.no_grad()
    def policy_value_uncertainty(self, envs: Sequence[O.Game], actions: Sequence[Sequence[int]]):
        """`impl Agent for Net` (net6_simhash.rs:259-324): per position (logits of the legal moves, value,
        variance)."""
        n = self.n
        xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(self.cin, n, n) for g in envs]))
        policy, values, ube = self.forward(xs)
        policy = policy.reshape(len(envs), -1)
        out_logits: List[np.ndarray] = []
        for i, acts in enumerate(actions):
            idx = torch.tensor([O.move_index(n, a) for a in acts], dtype=torch.long)
            out_logits.append(policy[i, idx].numpy().astype(np.float32))
        if self.has_rnd:
            local = self.normalized_rnd(xs)
        else:
            local = torch.tensor([0.0 if int(i) in self.simhash_set else MAXIMUM_VARIANCE for i in self.get_indices(xs)])
        unc = torch.clamp(torch.maximum(torch.exp(ube.view(-1)), local), 0.0, MAXIMUM_VARIANCE)
        return out_logits, values.view(-1).numpy().astype(np.float32), unc.numpy().astype(np.float32)

    def as_array_agent(self, device: str = "cpu"):
        """The same agent for bulk use (oracle.array_agent): planes and move indices come from the batched C
        helpers, the f32 forward runs on `device` (on "cuda" with TF32 off: plain f32 like the CPU path) and the
        legal logits are gathered with one index op.  Empty SimHash set only (variance = 4.0 like a fresh network)."""
        assert not self.simhash_set
        dev = torch.device(device)
        if dev.type == "cuda":
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        net = self.to(dev)
        L = O.lib()
        n = self.n

        @torch.no_grad()
        def fn(games, batch, act, na):
            xs = torch.from_numpy(O.game_repr_batch(games, batch)).to(dev)
            idx = np.zeros(act.shape, dtype=np.int32)
            L.tk_move_index_batch(n, act.ctypes.data, na.ctypes.data, act.shape[1], batch, idx.ctypes.data)
            policy, values, ube = net.forward(xs)
            lg = torch.gather(policy.reshape(batch, -1), 1, torch.from_numpy(idx).to(dev).long())
            unc = torch.clamp(torch.maximum(torch.exp(ube.view(-1)), torch.full((batch,), MAXIMUM_VARIANCE, device=dev)),
                              0.0, MAXIMUM_VARIANCE)
            return lg.float().cpu().numpy(), values.view(-1).float().cpu().numpy(), unc.float().cpu().numpy()

        return O.array_agent(fn)

    def as_oracle_agent(self):
        """Adapter to oracle.py_agent: (envs, actions) -> (logits, values, variances)."""
        return O.py_agent(lambda envs, acts: self.policy_value_uncertainty(envs, acts))
