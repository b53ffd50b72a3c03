"""CPU: the library's libtorch-free reader of the reference's model files (takzero_b200/csrc/model_file.cpp,
tz_read_model_file) -- `model_latest.ot` is a tch `VarStore::save` archive (takzero/src/network/mod.rs:16-18,
net6_simhash.rs:152-171), i.e. a libtorch module archive: ZIP + pickle.

Pinned by tests/golden/tiny_model.ot, written by libtorch's own `OutputArchive` the way tch does
(tests/golden/make_ot_fixture.cpp), and cross-checked against torch's readers on files written here."""
import json
import os

import numpy as np
import pytest
import torch

from takzero_b200 import capi, network, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reads_the_libtorch_written_fixture_bit_exactly():
    path = os.path.join(GOLDEN, "tiny_model.ot")
    want = json.load(open(os.path.join(GOLDEN, "tiny_model_expected.json")))
    got = network.read_model_file(path, stored_names=True)
    assert list(got) == list(want)  # file order
    for name, w in want.items():
        assert list(got[name].shape) == w["shape"]
        assert got[name].view(np.uint32).reshape(-1).tolist() == w["bits"]
    # and against torch's own reader, live
    live = dict(torch.jit.load(path).named_buffers())
    for name in want:
        assert np.array_equal(got[name], live[name].numpy())


def test_tch_variable_names_map_to_the_library_names():
    got = network.read_model_file(os.path.join(GOLDEN, "tiny_model.ot"))
    assert list(got) == ["core.input_conv2d.weight", "core.batch_norm.running_mean",
                         "core.res_block_0.0.conv2d.weight", "core.res_block_0.1.conv2d.weight", "value.linear.bias"]


@pytest.mark.parametrize("n", [4, 6])
def test_full_network_archive_round_trip(tmp_path, n):
    """A whole reference-shaped network under tch's names (second SmallBlock suffixed `__K`, residual.rs:52-54)
    comes back under the names tz_set_weights documents, bit for bit."""
    w = weights.random_init(n, seed=5, blocks=3)
    w["simhash_matrix"] = np.random.default_rng(1).standard_normal((weights.channels(n)[0] * n * n, 32)).astype(np.float32)
    path = tmp_path / "model_latest.ot"
    weights.save_ot(str(path), w)
    stored = network.read_model_file(str(path), stored_names=True)
    assert "core.res_block_1.conv2d.weight" in stored
    assert any(k.startswith("core.res_block_1.conv2d.weight__") for k in stored)
    assert not any(".0.conv2d" in k or ".1.conv2d" in k for k in stored)
    got = network.read_model_file(str(path))
    assert set(got) == set(w)
    for k in w:
        assert got[k].shape == w[k].shape and np.array_equal(got[k], w[k]), k


def test_torch_save_state_dict_dtypes_and_strides(tmp_path):
    sd = {
        "a.weight": torch.randn(4, 6).t(),                    # non-contiguous view
        "b": torch.randn(3, 5)[1:, ::2],                      # storage offset + strides
        "half": torch.randn(7).half(),
        "bf16": torch.randn(2, 2).bfloat16(),
        "f64": torch.randn(3).double(),
        "steps": torch.arange(5),
        "scalar": torch.tensor(2.5),
        "empty": torch.zeros(0, 3),
    }
    path = tmp_path / "sd.pt"
    torch.save(sd, str(path))
    got = network.read_model_file(str(path))
    assert set(got) == set(sd)
    for k, t in sd.items():
        assert got[k].shape == tuple(t.shape), k
        assert np.array_equal(got[k], t.float().numpy()), k


def test_nested_module_state_and_tzw(tmp_path):
    net = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 3), torch.nn.BatchNorm2d(3))
    torch.save(net.state_dict(), str(tmp_path / "m.pt"))
    got = network.read_model_file(str(tmp_path / "m.pt"))
    for k, t in net.state_dict().items():
        assert np.array_equal(got[k], t.float().numpy()), k
    w = weights.random_init(4, seed=2, blocks=1)
    weights.save_tzw(str(tmp_path / "w.tzw"), w)
    got = network.read_model_file(str(tmp_path / "w.tzw"))
    assert list(got) == list(w)
    for k in w:
        assert np.array_equal(got[k], w[k])


def test_bad_files_fail_with_a_message(tmp_path):
    with pytest.raises(capi.TakzeroError, match="cannot open"):
        network.read_model_file(str(tmp_path / "absent.ot"))
    (tmp_path / "junk.ot").write_bytes(b"not a model at all")
    with pytest.raises(capi.TakzeroError, match="neither"):
        network.read_model_file(str(tmp_path / "junk.ot"))
    blob = open(os.path.join(GOLDEN, "tiny_model.ot"), "rb").read()
    (tmp_path / "cut.ot").write_bytes(blob[: len(blob) // 2])  # what a reader sees while `learn` is still writing
    with pytest.raises(capi.TakzeroError):
        network.read_model_file(str(tmp_path / "cut.ot"))
    (tmp_path / "cut.tzw").write_bytes(b"TZW1" + (3).to_bytes(4, "little") + b"\x05\x00\x00\x00ab")
    with pytest.raises(capi.TakzeroError, match="truncated"):
        network.read_model_file(str(tmp_path / "cut.tzw"))


def test_safetensors_files(tmp_path):
    """tch's `VarStore::save` writes safetensors when the path ends in ".safetensors": u64 header length, JSON header,
    raw little-endian data.  Read by the same entry point, names mapped like the `.ot` ones."""
    from safetensors.numpy import save_file
    from safetensors.torch import save_file as save_torch

    w = weights.random_init(4, seed=3, blocks=2)
    save_file({k: np.ascontiguousarray(v) for k, v in weights.tch_names(w).items()}, str(tmp_path / "m.safetensors"),
              metadata={"format": "pt", "note": 'a "quoted" value, {braces} and [brackets]'})
    got = network.read_model_file(str(tmp_path / "m.safetensors"))
    assert set(got) == set(w)
    for k in w:
        assert got[k].shape == w[k].shape and np.array_equal(got[k], w[k]), k
    t = {"h": torch.randn(5).half(), "b": torch.randn(2, 3).bfloat16(), "d": torch.randn(3).double(),
         "i": torch.arange(4), "e": torch.zeros(0), "s": torch.tensor(1.5)}
    save_torch(t, str(tmp_path / "t.safetensors"))
    got = network.read_model_file(str(tmp_path / "t.safetensors"))
    assert set(got) == set(t)
    for k, v in t.items():
        assert got[k].shape == tuple(v.shape) and np.array_equal(got[k], v.float().numpy()), k
    blob = open(tmp_path / "m.safetensors", "rb").read()
    (tmp_path / "cut.safetensors").write_bytes(blob[: len(blob) - 1000])
    with pytest.raises(capi.TakzeroError, match="data_offsets"):
        network.read_model_file(str(tmp_path / "cut.safetensors"))
