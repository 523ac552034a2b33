"""CPU tests of the N > 1 path: slice rule, and a world_size-2 gloo run of shard -> compute ->
host-side gather, with the oracle standing in for the per-rank GPU call (tests only)."""
import os
import socket

import pytest

from pbc_b200.sharding import shard_bounds, shard_inputs


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 2, 7, 8, 9, 1 << 20, (1 << 18) + 777):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            per = (n + world - 1) // world
            assert all(hi - lo <= per for lo, hi in b)


def test_shard_inputs_never_split_a_product():
    k, l1, l2 = 16, 128, 128
    n_out = 10
    in1 = bytes(range(256)) * (n_out * k * l1 // 256)
    in2 = in1[::-1]
    seen = 0
    for r in range(4):
        a, b, m = shard_inputs(in1, in2, n_out, k, l1, l2, r, 4)
        assert len(a) == m * k * l1 and len(b) == m * k * l2
        assert a == in1[seen * k * l1:(seen + m) * k * l1]
        seen += m
    assert seen == n_out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, k, n_out, ret):
    import json
    import torch.distributed as dist
    from oracle import pbc_oracle as O
    from pbc_b200.params import PARAMS
    from pbc_b200.sharding import sharded_apply
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(root, "tests", "golden", name + ".json")))
    orc = O.pairing_from_param(PARAMS[name])
    P = b"".join(bytes.fromhex(x) for x in g["prod"]["P"][:k * n_out])
    Q = b"".join(bytes.fromhex(x) for x in g["prod"]["Q"][:k * n_out])

    def compute(a, b, m):      # stands in for Pairing.prod_apply on this rank's GPU
        return b"".join(O.prod_pairing_bytes(orc, [a[(i * k + j) * orc.g1_len:(i * k + j + 1) * orc.g1_len] for j in range(k)],
                                             [b[(i * k + j) * orc.g2_len:(i * k + j + 1) * orc.g2_len] for j in range(k)])
                        for i in range(m))

    out = sharded_apply(compute, P, Q, n_out, k, orc.g1_len, orc.g2_len, orc.gt_len, rank, world)
    if rank == 0:
        ret["out"] = out
        ret["want"] = b"".join(bytes.fromhex(x) for x in g["prod"]["e"][:n_out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["d159", "a"])
def test_two_rank_gloo_shard_compute_gather(name):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    k, n_out = (4, 3)          # 3 outputs over 2 ranks: ragged split (2 + 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, name, k, n_out, ret), nprocs=world, join=True)
    assert ret["out"] == ret["want"]
