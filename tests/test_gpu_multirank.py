"""GPU, two ranks on two GPUs of one node (skipped on a single-GPU box; run with `gpurun --gpus 2`): the library's own
NCCL communicator (csrc/comm.cu).  One weight generation broadcast from rank 0 must leave every rank with exactly the
weight set a local tz_set_weights of the same tensors produces (bit for bit) and with identical network outputs, the
swap must happen between moves, and tz_allreduce_sum must add up the per-rank counters.  Replaces the per-process
`Net::load(model_latest.ot)` of selfplay/src/main.rs:107 and the file-based sums of learn/src/main.rs:195-209."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from oracle import net_ref
    from oracle import oracle as O
    from takzero_b200 import capi, network
    from takzero_b200 import distributed as tzd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only: carries the 128-byte unique id
    try:
        n, hk, G = 4, 4, 64
        base, _ = tzd.shard(rank, world, G)
        m = capi.BatchedMCTS(n, hk, G, device=rank, game_base=base, arena_slots=1 << 14)
        tzd.init_comm(m, rank, world)
        nets = [net_ref.Net(n, seed=s, blocks=2, randomize_bn=True) for s in (1, 2)]
        games = [O.new_opening(n, hk, i % 8, i // 8 % 2) for i in range(G)]
        states = O.pack_games(games).view(capi.STATE_DTYPE).reshape(-1)
        actions = [O.possible_moves(g) for g in games]
        results = []
        for gen, net in enumerate(nets):
            network.broadcast_weights(m, net.tensors() if rank == 0 else None, root=0, res_blocks=2)
            if gen == 0:
                m.set_agent(capi.AGENT_NETWORK)
            got_set = network.weight_set(m)
            local = capi.BatchedMCTS(n, hk, G, device=rank, arena_slots=4096)
            network.set_weights(local, net.tensors())
            want_set = network.weight_set(local)
            want = network.evaluate(local, states, actions)
            local.close()
            got = network.evaluate(m, states, actions)
            # bytes 16..23 of the header hold the generation counter (2 here, 1 on the fresh local handle)
            same = bool(np.array_equal(got_set[:16], want_set[:16]) and np.array_equal(got_set[24:], want_set[24:]))
            same = same and all(np.array_equal(a, b) for a, b in zip(got[0], want[0])) and np.array_equal(got[1], want[1])
            results.append(same)
        # a search move between the generations still works, and the counters add up over the ranks
        m.new_openings(seed=5)
        m.gumbel_sequential_halving(None, 8, 48, None, seed=3)
        c = m.counters()
        total = network.allreduce_sum(m, [c.simulations, 1, rank])
        out[rank] = (results, int(c.simulations), total, network.weight_generation(m)[0], m.status())
        m.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_broadcast_weights_and_sum_counters():
    import torch.multiprocessing as mp

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        results, sims, total, generation, status = out[r]
        assert results == [True, True], f"rank {r}: broadcast weights differ from a local load"
        assert status == 0 and generation == 2
        assert total == [out[0][1] + out[1][1], 2, 1]
