"""GPU parity: encoding (bit-exact) and the 16-bit tcgen05 network vs the f32 PyTorch restatement.

Tolerances (stated per BASELINE.json north_star): planes bit-exact; policy logits / value within
|delta| <= 1e-2 of the f32 reference on random-init weights; argmax agreement >= 99%."""
import numpy as np
import pytest
import torch

from oracle import net_ref
from oracle import oracle as O
from takzero_b200 import capi, network

from helpers import games_to_states, random_playout_states

pytestmark = pytest.mark.gpu

TOL = 1e-2


def sample_positions(n, half_komi, count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        ps = random_playout_states(n, half_komi, int(rng.integers(1 << 30)), keep_terminal=False)
        for i in rng.choice(len(ps), size=min(4, len(ps)), replace=False):
            out.append(ps[int(i)])
    return out[:count]


@pytest.mark.parametrize("n,half_komi", [(3, 0), (4, 4), (5, 4), (6, 4)])
def test_encode_planes_bit_exact(n, half_komi):
    games = sample_positions(n, half_komi, 200, n)
    m = capi.BatchedMCTS(n, half_komi, 4, arena_slots=4096)
    got = network.encode_planes(m, games_to_states(games))
    want = np.stack([O.game_repr(g) for g in games]).reshape(got.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    m.close()


def test_encode_golden_vector():
    """repr.rs:303-360 `complicated_position` through the CUDA encoder."""
    g = O.from_tps(5, 4, "x2,1221,x,1S/2,2C,2,1,x/x,212,21C,2S,2/2211S,2,21,1,1/x2,221S,2,x 2 23")
    m = capi.BatchedMCTS(5, 4, 2, arena_slots=4096)
    got = network.encode_planes(m, games_to_states([g]))[0]
    assert np.array_equal(got.reshape(-1).view(np.uint32), O.game_repr(g).view(np.uint32))
    assert got[-1, 0, 0] == np.float32(-3.0) / np.float32(25.0)  # FCD plane
    m.close()


@pytest.mark.parametrize("n,half_komi", [(4, 4), (5, 4), (6, 4)])
@pytest.mark.parametrize("dtype", [network.DTYPE_F16, network.DTYPE_BF16])
def test_first_convolution_input_planes_are_game_repr_in_16_bits(n, half_komi, dtype):
    """The network never stores its input planes: the first convolution's A producer encodes them from the queued
    positions (conv_tcgen05.cuh).  The hook runs the same device function: every value must be game_repr's
    (repr.rs:169-228) rounded to the network's 16-bit type, channels beyond C zero."""
    count = 96
    games = sample_positions(n, half_komi, count, 7 * n)
    actions = [O.possible_moves(g) for g in games]
    m = capi.BatchedMCTS(n, half_komi, count, arena_slots=4096)
    network.set_weights(m, net_ref.Net(n, seed=1, blocks=1).tensors(), dtype)
    network.evaluate(m, games_to_states(games), actions)
    got = network.debug_activations(m, 2, count)  # [count, N*N, 64]
    want = np.stack([O.game_repr(g) for g in games]).reshape(count, -1, n * n).transpose(0, 2, 1)  # [count, N*N, C]
    C_ = want.shape[2]
    if dtype == network.DTYPE_F16:
        want16 = want.astype(np.float16).astype(np.float32)
    else:
        want16 = torch.from_numpy(np.ascontiguousarray(want)).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(got[:, :, :C_].view(np.uint32), want16.view(np.uint32))
    assert (got[:, :, C_:] == 0).all()
    m.close()


def check_network(n, half_komi, count, blocks, seed, randomize_bn, dtype=network.DTYPE_BF16, tol=TOL):
    ref = net_ref.Net(n, seed=seed, blocks=blocks, randomize_bn=randomize_bn)
    games = sample_positions(n, half_komi, count, seed)
    actions = [O.possible_moves(g) for g in games]
    m = capi.BatchedMCTS(n, half_komi, count, arena_slots=4096)
    network.set_weights(m, ref.tensors(), dtype)
    logits, values, variances = network.evaluate(m, games_to_states(games), actions)
    want_logits, want_values, want_var = ref.policy_value_uncertainty(games, actions)
    worst = 0.0
    agree = 0
    for i in range(count):
        assert logits[i].shape == want_logits[i].shape
        worst = max(worst, float(np.abs(logits[i] - want_logits[i]).max()))
        agree += int(np.argmax(logits[i]) == np.argmax(want_logits[i]))
    dv = float(np.abs(values - want_values).max())
    print(f"n={n} blocks={blocks}: max |dlogit| {worst:.4g}, max |dvalue| {dv:.4g}, argmax agreement {agree}/{count}")
    assert worst <= tol, f"policy logits differ by {worst}"
    assert dv <= tol, f"values differ by {dv}"
    assert np.array_equal(variances, want_var)  # 4.0 everywhere with an empty SimHash set
    assert agree >= 0.99 * count
    m.close()


def test_single_conv_layer_matches_torch():
    """First convolution only (input planes -> 256 channels), compared element-wise."""
    n, hk, count = 6, 4, 40
    ref = net_ref.Net(n, seed=5, blocks=1, randomize_bn=True)
    games = sample_positions(n, hk, count, 5)
    actions = [O.possible_moves(g) for g in games]
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(m, ref.tensors())
    network.debug_layer_limit(m, 1)
    network.evaluate(m, games_to_states(games), actions)
    got = network.debug_activations(m, 0, count)  # [count, 36, 256]
    xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(ref.cin, n, n) for g in games]))
    with torch.no_grad():
        want = torch.relu(ref.core.batch_norm(ref.core.input_conv2d(xs)))
    want = want.permute(0, 2, 3, 1).reshape(count, n * n, 256).numpy()
    err = np.abs(got - want).max()
    print("first conv max abs err", err, "ref max", np.abs(want).max())
    assert err <= 2e-2 * max(1.0, float(np.abs(want).max()))
    network.debug_layer_limit(m, 3)
    network.evaluate(m, games_to_states(games), actions)
    got = network.debug_activations(m, 0, count)
    with torch.no_grad():
        want = ref.core.res_block_0(torch.relu(ref.core.batch_norm(ref.core.input_conv2d(xs))))
    want = want.permute(0, 2, 3, 1).reshape(count, n * n, 256).numpy()
    err = np.abs(got - want).max()
    print("block 0 max abs err", err, "ref max", np.abs(want).max())
    assert err <= 3e-2 * max(1.0, float(np.abs(want).max()))
    m.close()


def test_network_small_4x4():
    check_network(4, 4, 64, 2, 11, True)


def test_network_full_6x6():
    check_network(6, 4, 96, 16, 123, False)


def test_network_full_4x4():
    check_network(4, 4, 128, 16, 123, False)


def test_network_full_5x5():
    check_network(5, 4, 48, 20, 123, False)


def test_network_fp16_mode_is_tighter():
    """TZ_DTYPE_F16: same kernels on IEEE half; the error against the f32 reference is ~8x smaller."""
    check_network(6, 4, 96, 16, 123, False, dtype=network.DTYPE_F16, tol=1.5e-3)
    check_network(4, 4, 64, 2, 11, True, dtype=network.DTYPE_F16, tol=1.5e-3)


def test_search_with_device_network_matches_injected_outputs():
    """The search driven by the device network == the search with that network's own outputs
    injected through the host callback (the tree code is identical; only the data path differs)."""
    n, hk, G = 4, 4, 16
    ref = net_ref.Net(n, seed=3, blocks=2)
    a = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 14)
    b = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 14)
    evaluator = capi.BatchedMCTS(n, hk, G, arena_slots=4096)
    for h in (a, b, evaluator):
        network.set_weights(h, ref.tensors())
    a.new_openings(seed=1)
    b.new_openings(seed=1)
    a.set_agent(capi.AGENT_NETWORK)

    import ctypes as C

    def cb(_ctx, batch, envs, actions, n_actions, stride, logits, values, variances):
        st = np.ctypeslib.as_array(C.cast(envs, C.POINTER(C.c_uint8)), shape=(batch * 384,)).view(capi.STATE_DTYPE)
        act = np.ctypeslib.as_array(actions, shape=(batch, stride))
        na = np.ctypeslib.as_array(n_actions, shape=(batch,))
        lg, v, u = network.evaluate(evaluator, st.copy(), [list(act[i, : na[i]]) for i in range(batch)])
        lo = np.ctypeslib.as_array(logits, shape=(batch, stride))
        for i in range(batch):
            lo[i, : na[i]] = lg[i]
        np.ctypeslib.as_array(values, shape=(batch,))[:] = v
        np.ctypeslib.as_array(variances, shape=(batch,))[:] = u

    b.set_agent(capi.AGENT_HOST, cb)
    gum = np.random.default_rng(0).gumbel(size=(G, a.move_stride)).astype(np.float32)
    ma = a.gumbel_sequential_halving(None, 8, 48, gum)
    mb = b.gumbel_sequential_halving(None, 8, 48, gum)
    assert np.array_equal(ma, mb)
    ta, tb = a.root_children(), b.root_children()
    assert np.array_equal(ta["visits"], tb["visits"])
    assert np.array_equal(ta["eval_bits"], tb["eval_bits"])
    for h in (a, b, evaluator):
        h.close()


def test_simhash_indices_and_uncertainty():
    """SimHash novelty (net6_simhash.rs:203-256,309-317): hash index per position, set lookup and the
    uncertainty combine clamp(max(exp(ube), local), 0, 4)."""
    n, hk, count = 6, 4, 64
    ref = net_ref.Net(n, seed=9, blocks=1)
    games = sample_positions(n, hk, count, 9)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(m, ref.tensors())
    network.set_simhash(m, ref.simhash_matrix.numpy())
    got = network.simhash_indices(m, states)
    xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(ref.cin, n, n) for g in games]))
    dots = ref.simhash_dots(xs).numpy()
    want = ref.get_indices(xs)
    # bits may only differ where the f32 dot product is within rounding of zero (summation order)
    diff = got ^ want
    for i in range(count):
        for b in range(32):
            if (int(diff[i]) >> b) & 1:
                assert abs(dots[i, b]) < 1e-3, f"position {i} bit {b}: dot {dots[i, b]}"
    assert (diff == 0).mean() > 0.9
    # half of the positions are "seen": their local uncertainty drops to 0, so variance = clamp(exp(ube))
    ref.simhash_set = {int(x) for x in got[::2]}
    network.set_simhash(m, ref.simhash_matrix.numpy(), ref.bitset_bytes())
    _, _, variances = network.evaluate(m, states, actions)
    with torch.no_grad():
        ube = ref.ube(ref.core(xs)).view(-1)
    seen = np.array([int(x) in ref.simhash_set for x in got])
    want_var = np.where(seen, np.clip(np.exp(ube.numpy()), 0.0, 4.0), 4.0)
    assert seen.sum() >= count // 2
    assert np.abs(variances - want_var).max() <= 2e-2
    assert (variances[~seen] == 4.0).all()
    m.close()


def test_update_counts_marks_positions_seen_and_survives_a_save(tmp_path):
    """The reference's `counts_work` and `saving_works` (net6_simhash.rs:368-430): positions five random plies past an
    opening are all novel for a fresh network (local uncertainty 4.0); after update_counts every one of them is seen
    (uncertainty strictly lower); the set written like Net::save's bitvec.bin and loaded into another handle says the
    same.  The set's image has exactly the bits of `get_indices` of those positions."""
    from takzero_b200 import weights

    n, hk, count = 6, 4, 128
    ref = net_ref.Net(n, seed=456, blocks=1)
    games = sample_positions(n, hk, count, 456)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    a = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(a, ref.tensors())
    network.set_simhash(a, ref.simhash_matrix.numpy())
    before = network.evaluate(a, states, actions)[2]
    assert (before == 4.0).all()
    network.update_counts(a, states)
    after = network.evaluate(a, states, actions)[2]
    # forward_hash went from 4.0 to 0.0 for every position, so what is left is clamp(exp(ube), 0, 4)
    xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(ref.cin, n, n) for g in games]))
    with torch.no_grad():
        ube = ref.ube(ref.core(xs)).view(-1).numpy()
    assert np.abs(after - np.clip(np.exp(ube), 0.0, 4.0)).max() <= 2e-2
    assert (before > after)[np.exp(ube) < 3.9].all() and (np.exp(ube) < 3.9).mean() > 0.9
    idx = network.simhash_indices(a, states)
    image = network.read_novelty_set(a)
    want = np.zeros(1 << 29, dtype=np.uint8)
    np.bitwise_or.at(want, idx >> 3, (1 << (idx & 7)).astype(np.uint8))
    assert np.array_equal(image, want)
    # a second update with the same positions changes nothing (the set only grows)
    network.update_counts(a, states[: count // 2])
    assert np.array_equal(network.read_novelty_set(a), want)

    tensors = dict(ref.tensors())
    tensors["simhash_matrix"] = ref.simhash_matrix.numpy()
    weights.save_ot(str(tmp_path / "delete-me.ot"), tensors)
    image.tofile(str(tmp_path / "bitvec.bin"))
    b = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.load_model(b, str(tmp_path / "delete-me.ot"))
    assert np.array_equal(network.evaluate(b, states, actions)[2], after)
    for h in (a, b):
        h.close()


@pytest.mark.parametrize("n,hk,count", [(4, 4, 701), (6, 4, 333), (5, 4, 97)])
def test_chunked_fused_launch_equals_one_chunk_and_per_layer_launches(n, hk, count):
    """The network body runs as one persistent launch whose positions are cut into chunks that walk through all
    layers on small L2-resident activation sets (conv_tcgen05.cuh).  Chunk and tile placement must not change a
    single bit: many small chunks (fewer tiles than CTA pairs, partial last tile), one chunk, and one launch per
    layer give identical logits / values, and they agree with the f32 reference."""
    ref = net_ref.Net(n, seed=3, blocks=3, randomize_bn=True)
    games = sample_positions(n, hk, count, 77)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    outs = {}
    for name, mode in (("chunks", {"chunk_min_tiles": 3}), ("one", {"chunk_min_tiles": 0}),
                       ("layers", {"per_layer_launches": True})):
        m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
        network.debug_network_mode(m, **mode)
        network.set_weights(m, ref.tensors())  # the mode is read when the weights are set
        outs[name] = network.evaluate(m, states, actions)
        for _ in range(4 if name == "chunks" else 0):  # the cross-CTA hand-offs are timing dependent: repeat
            again = network.evaluate(m, states, actions)
            assert all(np.array_equal(a, b) for a, b in zip(outs[name][0], again[0]))
            assert np.array_equal(outs[name][1], again[1])
        assert m.status() == 0
        m.close()
    for name in ("one", "layers"):
        for a, b in zip(outs["chunks"][0], outs[name][0]):
            assert np.array_equal(a, b), name
        assert np.array_equal(outs["chunks"][1], outs[name][1]), name
    want_logits, want_values, _ = ref.policy_value_uncertainty(games, actions)
    assert max(float(np.abs(a - b).max()) for a, b in zip(outs["chunks"][0], want_logits)) <= TOL
    assert float(np.abs(outs["chunks"][1] - want_values).max()) <= TOL


@pytest.mark.parametrize("count", [1, 8, 9, 130, 1024, 1184, 1185])
def test_local_chain_of_4x4_small_batches_is_bit_identical(count):
    """4x4: 16 squares divide the 128 rows of a CTA, so with at most one tile per CTA pair (<= 1184 positions) the fused
    launch keeps every tile's activations in shared memory from layer to layer (conv_tcgen05.cuh "local chain").  Same
    bits as one launch per layer through global memory, for a single position, exactly one CTA, a ragged last tile, the
    BASELINE configs[1] size, the largest local size and the first size that is not local any more."""
    n, hk = 4, 4
    ref = net_ref.Net(n, seed=8, blocks=3, randomize_bn=True)
    games = sample_positions(n, hk, min(count, 256), 99)
    games = [games[i % len(games)] for i in range(count)]
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    outs = []
    for per_layer in (False, True):
        m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
        network.debug_network_mode(m, per_layer_launches=per_layer)
        network.set_weights(m, ref.tensors())
        outs.append(network.evaluate(m, states, actions))
        if not per_layer:
            again = network.evaluate(m, states, actions)
            assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], again[0])) and np.array_equal(outs[0][1], again[1])
        assert m.status() == 0
        m.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], outs[1][0]))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    want_logits, want_values, _ = ref.policy_value_uncertainty(games[:64], actions[:64])
    assert max(float(np.abs(a - b).max()) for a, b in zip(outs[0][0][:64], want_logits)) <= TOL
    assert float(np.abs(outs[0][1][:64] - want_values).max()) <= TOL


@pytest.mark.parametrize("n,count", [(6, 1), (6, 3), (6, 4), (6, 128), (6, 444), (6, 445), (5, 5), (5, 6), (5, 128),
                                     (5, 740), (5, 741)])
def test_packed_local_chain_of_5x5_and_6x6_small_batches_is_bit_identical(n, count):
    """5x5 / 6x6: 25 / 36 squares do not divide the 128 rows of a CTA.  With at most one tile per CTA (<= 740 / 444
    positions on 148 SMs) the fused launch packs 5 / 3 whole positions into every CTA tile (the rest of its rows are
    dead), so no position straddles two CTAs and the shared-memory local chain applies.  Same bits as one launch per
    layer on dense rows: a single position, exactly one tile, one position more, the 128 leaves of tei's batches, the
    largest packed size and the first one that is dense again."""
    hk = 4
    ref = net_ref.Net(n, seed=8, blocks=3, randomize_bn=True)
    games = sample_positions(n, hk, min(count, 200), 99)
    games = [games[i % len(games)] for i in range(count)]
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    outs = []
    for per_layer in (False, True):
        m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
        network.debug_network_mode(m, per_layer_launches=per_layer)
        network.set_weights(m, ref.tensors())
        outs.append(network.evaluate(m, states, actions))
        if not per_layer:
            again = network.evaluate(m, states, actions)
            assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], again[0])) and np.array_equal(outs[0][1], again[1])
        assert m.status() == 0
        m.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], outs[1][0]))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    want_logits, want_values, _ = ref.policy_value_uncertainty(games[:32], actions[:32])
    assert max(float(np.abs(a - b).max()) for a, b in zip(outs[0][0][:32], want_logits)) <= TOL
    assert float(np.abs(outs[0][1][:32] - want_values).max()) <= TOL


def test_watchdog_turns_a_stalled_dependency_into_a_status_bit():
    """The fused launch waits on other CTA pairs' progress counters.  With the test hook that makes pair 0 withhold
    its tiles, the dependent pairs must not spin forever: the watchdog raises TZ_STATUS_NETWORK_STALL (256), the
    launch ends, and the handle reports the error instead of hanging the GPU."""
    import time

    n, hk, count = 5, 4, 800  # 5x5, more tiles than CTAs: dense rows, positions straddle tiles, the pairs depend on each other
    ref = net_ref.Net(n, seed=2, blocks=2)
    games = sample_positions(n, hk, count, 5)
    actions = [O.possible_moves(g) for g in games]
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(m, ref.tensors())
    good = network.evaluate(m, games_to_states(games), actions)
    assert m.status() == 0
    network.debug_network_mode(m, drop_progress=True)
    t0 = time.perf_counter()
    try:
        network.evaluate(m, games_to_states(games), actions)
    except capi.TakzeroError:
        pass
    assert time.perf_counter() - t0 < 60.0
    assert m.status() & 256
    m.close()
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)  # a fresh handle works as before
    network.set_weights(m, ref.tensors())
    again = network.evaluate(m, games_to_states(games), actions)
    assert m.status() == 0 and np.array_equal(good[1], again[1])
    m.close()


def test_load_model_from_tch_archive_with_bitvec_sidecar(tmp_path):
    """Net::load (network/mod.rs:20-27, net6_simhash.rs:164-181) from the reference's own files: a tch-named
    `model_latest.ot` plus the `bitvec.bin` SimHash set beside it give exactly the outputs of tz_set_weights +
    tz_set_simhash with the same tensors (bit for bit), in both weight dtypes."""
    from takzero_b200 import weights

    n, hk, count = 6, 4, 48
    ref = net_ref.Net(n, seed=13, blocks=2, randomize_bn=True)
    games = sample_positions(n, hk, count, 13)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    a = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(a, ref.tensors())
    network.set_simhash(a, ref.simhash_matrix.numpy())
    ref.simhash_set = {int(x) for x in network.simhash_indices(a, states)[::3]}
    bits = ref.bitset_bytes()
    network.set_simhash(a, ref.simhash_matrix.numpy(), bits)
    want = network.evaluate(a, states, actions)
    assert (want[2] < 4.0).sum() >= count // 3

    tensors = dict(ref.tensors())
    tensors["simhash_matrix"] = ref.simhash_matrix.numpy()
    weights.save_ot(str(tmp_path / "model_latest.ot"), tensors)
    np.asarray(bits, dtype=np.uint8).tofile(str(tmp_path / "bitvec.bin"))
    b = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.load_model(b, str(tmp_path / "model_latest.ot"))
    got = network.evaluate(b, states, actions)
    for x, y in zip(want[0], got[0]):
        assert np.array_equal(x, y)
    assert np.array_equal(want[1], got[1]) and np.array_equal(want[2], got[2])
    # without the sidecar: an error like Net::load's (net6_simhash.rs:164-181) that leaves the loaded model in place,
    # unless the caller asks for the empty set of a fresh network (every local uncertainty is then 4.0)
    (tmp_path / "bitvec.bin").unlink()
    with pytest.raises(capi.TakzeroError, match="bitvec.bin is missing"):
        network.load_model(b, str(tmp_path / "model_latest.ot"))
    assert np.array_equal(network.evaluate(b, states, actions)[2], want[2])
    network.load_model(b, str(tmp_path / "model_latest.ot"), allow_missing_set=True)
    assert (network.evaluate(b, states, actions)[2] == 4.0).all()
    # a file that lacks a tensor is refused with its name, and the previous model stays usable
    tensors.pop("policy.conv2d.bias")
    weights.save_ot(str(tmp_path / "broken.ot"), tensors)
    with pytest.raises(capi.TakzeroError, match="policy.conv2d.bias"):
        network.load_model(b, str(tmp_path / "broken.ot"), allow_missing_set=True)
    again = network.evaluate(b, states, actions)
    assert all(np.array_equal(x, y) for x, y in zip(want[0], again[0])) and np.array_equal(want[1], again[1])
    for h in (a, b):
        h.close()


@pytest.mark.parametrize("n,hk", [(4, 4), (6, 4)])
def test_lcghash_indices_bit_exact_and_uncertainty(n, hk, tmp_path):
    """LCG-hash novelty of the reference's net4_lcghash.rs:203-261: an integer hash of the planes, so the index is
    bit-exact against the tensor-op restatement; the set lookup then works like SimHash's.  Also through
    tz_load_model (`lcghash_init` in the archive + bitvec.bin)."""
    from takzero_b200 import weights

    count = 96
    ref = net_ref.Net(n, seed=4, blocks=1)
    games = sample_positions(n, hk, count, 41)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    init = torch.empty(ref.cin, n, n).uniform_(-100.0, 100.0, generator=torch.Generator().manual_seed(5))
    xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(ref.cin, n, n) for g in games]))
    want = net_ref.Net.lcghash_indices(xs, init)
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.set_weights(m, ref.tensors())
    network.set_lcghash(m, init.numpy())
    got = network.lcghash_indices(m, states)
    assert np.array_equal(got, want)
    assert len(set(got.tolist())) > count // 2  # it does hash
    # mark every other position as seen: its local uncertainty drops to 0
    ref.simhash_set = {int(x) for x in got[::2]}
    bits = ref.bitset_bytes()
    network.set_lcghash(m, init.numpy(), bits)
    _, _, variances = network.evaluate(m, states, actions)
    with torch.no_grad():
        ube = ref.ube(ref.core(xs)).view(-1).numpy()
    seen = np.array([int(x) in ref.simhash_set for x in got])
    want_var = np.where(seen, np.clip(np.exp(ube), 0.0, 4.0), 4.0)
    assert np.abs(variances - want_var).max() <= 2e-2 and (variances[~seen] == 4.0).all()
    # the same through the model file
    tensors = dict(ref.tensors())
    tensors["lcghash_init"] = init.numpy()
    weights.save_ot(str(tmp_path / "model_latest.ot"), tensors)
    np.asarray(bits, dtype=np.uint8).tofile(str(tmp_path / "bitvec.bin"))
    b = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.load_model(b, str(tmp_path / "model_latest.ot"))
    assert np.array_equal(network.lcghash_indices(b, states), want)
    assert np.array_equal(network.evaluate(b, states, actions)[2], variances)
    # update_counts with this hash (net4_lcghash.rs has the same `update_counts`): the other half becomes seen too,
    # on top of the set that came with the file
    network.update_counts(b, states[1::2])
    after = network.evaluate(b, states, actions)[2]
    assert np.abs(after - np.clip(np.exp(ube), 0.0, 4.0)).max() <= 2e-2
    image = network.read_novelty_set(b)
    full = np.zeros(1 << 29, dtype=np.uint8)
    np.bitwise_or.at(full, got >> 3, (1 << (got & 7)).astype(np.uint8))
    assert np.array_equal(image, full)
    for h in (m, b):
        h.close()


def test_model_reload_replaces_weights_and_keeps_buffers():
    """Net::load before every move (selfplay/src/main.rs:107): a second tz_set_weights takes effect for the
    next evaluation (activation buffers are reused, only the folded weights change)."""
    n, hk, count = 4, 4, 32
    games = sample_positions(n, hk, count, 21)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    m = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    outs = []
    for seed in (1, 2, 1):
        ref = net_ref.Net(n, seed=seed, blocks=2)
        network.set_weights(m, ref.tensors())
        logits, values, _ = network.evaluate(m, states, actions)
        want_logits, want_values, _ = ref.policy_value_uncertainty(games, actions)
        assert max(float(np.abs(a - b).max()) for a, b in zip(logits, want_logits)) <= TOL
        assert float(np.abs(values - want_values).max()) <= TOL
        outs.append(np.concatenate(logits))
    assert np.array_equal(outs[0], outs[2]) and not np.array_equal(outs[0], outs[1])
    m.set_agent(capi.AGENT_NETWORK)
    m.new_openings(seed=1)
    m.gumbel_sequential_halving(None, 8, 48, None, seed=1)
    assert m.status() == 0
    m.close()


def test_rnd_local_uncertainty_of_the_5x5_network(tmp_path):
    """net5.rs:120-146,193-211,271-277: the 5x5 network's local uncertainty is random network distillation -- two MLPs
    (800 -> 1024 -> 1024 -> 512) on the planes divided by the sum of their squares, sum of squared differences,
    normalized with `min` / `max`, times 4, combined as clamp(max(exp(ube), rnd), 0, 4).  Device (TF32 tensor-op GEMMs)
    against the f32 restatement; the normalization is set so that the estimates spread over (0, 4)."""
    from takzero_b200 import weights

    n, hk, count = 5, 4, 96
    ref = net_ref.Net(n, seed=31, blocks=2, rnd=True)
    games = sample_positions(n, hk, count, 31)
    actions = [O.possible_moves(g) for g in games]
    states = games_to_states(games)
    xs = torch.from_numpy(np.stack([O.game_repr(g).reshape(ref.cin, n, n) for g in games]))
    with torch.no_grad():
        raw = (ref.rnd_learning(xs) - ref.rnd_target(xs)).square().sum(dim=1)
        # make exp(ube) small so that the RND term decides, and spread it over the range
        ref.ube.linear.bias.fill_(-6.0)
        ref.min.fill_(float(raw.min()) - 0.1 * float(raw.max() - raw.min()))
        ref.max.fill_(float(raw.max()) + 0.1 * float(raw.max() - raw.min()))
    _, _, want = ref.policy_value_uncertainty(games, actions)
    assert want.min() > 0.05 and want.max() < 3.95 and want.std() > 0.3
    m = capi.BatchedMCTS(n, hk, count, arena_slots=1 << 13)
    network.set_weights(m, ref.tensors())
    logits, values, got = network.evaluate(m, states, actions)
    print("RND uncertainty: max abs err", float(np.abs(got - want).max()), "range", float(want.min()), float(want.max()))
    assert np.abs(got - want).max() <= 2e-2
    # the search uses it: with beta > 0 the device-network search runs and the root's children carry these std devs
    m.set_agent(capi.AGENT_NETWORK)
    m.set_positions(states)
    m.gumbel_sequential_halving(np.full(count, 0.25, np.float32), 8, 24, None, seed=2)
    assert m.status() == 0
    std = m.root_children()["std_dev"]
    assert np.isfinite(std).all() and std.max() <= 2.0 + 1e-6 and std.max() > 0.2
    # through the model file (tch names) as well
    weights.save_ot(str(tmp_path / "model_latest.ot"), ref.tensors())
    b = capi.BatchedMCTS(n, hk, count, arena_slots=4096)
    network.load_model(b, str(tmp_path / "model_latest.ot"))
    assert np.array_equal(network.evaluate(b, states, actions)[2], got)
    # a model without the estimator switches it off again: 4.0 everywhere
    network.set_weights(b, net_ref.Net(n, seed=31, blocks=2).tensors())
    assert (network.evaluate(b, states, actions)[2] == 4.0).all()
    for h in (m, b):
        h.close()
