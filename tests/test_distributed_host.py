"""CPU tests of the multi-GPU host logic (no kernels): start level and splitters from the global
prefix histogram, result merging, and the communicator plumbing over gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cloud(kind, n, seed, **kw):
    from schwarzwald_b200 import synth
    return synth.generate(kind, n, seed, device="cpu", **kw).numpy()


def _bins_from_keys(keys):
    return np.bincount((keys >> np.uint64(45)).astype(np.int64), minlength=262144).astype(np.uint32)


@pytest.mark.parametrize("kind,n,conc", [("uniform", 1_200_000, 8), ("terrain", 900_000, 4), ("skewed", 700_000, 2),
                                          ("uniform", 50_000, 96)])
def test_start_level_from_global_histogram_matches_oracle(port_oracle, kind, n, conc):
    """swgpu_estimate_start_level on the summed level-5 histogram == the oracle's
    estimate_start_node_level_in_octree (TilingAlgorithms.cpp:1473-1535) on the sorted batch."""
    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import distributed
    xyz = _cloud(kind, n, 7, side_m=300.0)
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    params = sworacle.make_params("RANDOM_GRID", "FAST", spacing, bmin, bmax, concurrency=conc)
    want = port_oracle.tile(params, xyz)
    # the histogram is a sum over ranks: split the keys arbitrarily and add the parts
    keys = want.keys
    parts = np.array_split(keys, 3)
    bins = sum(_bins_from_keys(p).astype(np.uint64) for p in parts).astype(np.uint32)
    assert distributed.estimate_start_level(bins, conc) == want.start_level


def test_splitters_are_snapped_balanced_and_monotone():
    from schwarzwald_b200 import distributed
    rng = np.random.default_rng(1)
    bins = np.zeros(262144, np.uint32)
    occupied = rng.choice(262144, 5000, replace=False)
    bins[occupied] = rng.integers(1, 4000, 5000)
    total = int(bins.sum())
    for world in (1, 2, 4, 8, 16):
        for levels in (1, 3, 6):
            fp = distributed.choose_splitters(bins, world, levels)
            assert fp[0] == 0 and fp[-1] == 262144 and len(fp) == world + 1
            assert (np.diff(fp.astype(np.int64)) >= 0).all()
            group = 8 ** (6 - levels)
            assert (fp % group == 0).all(), "splitters must sit on shard-level subtree boundaries"
            if levels == 6:
                cum = np.concatenate([[0], np.cumsum(bins.astype(np.int64))])
                per_rank = np.diff(cum[fp])
                assert per_rank.sum() == total
                assert per_rank.max() <= total / world + 2 * bins.max()


def test_splitters_with_all_points_in_one_subtree():
    """Skew: one prefix holds everything -> one rank gets all points, the others none (no crash)."""
    from schwarzwald_b200 import distributed
    bins = np.zeros(262144, np.uint32)
    bins[777] = 123456
    fp = distributed.choose_splitters(bins, 4, 6)
    cum = np.concatenate([[0], np.cumsum(bins.astype(np.int64))])
    per_rank = np.diff(cum[fp])
    assert per_rank.sum() == 123456 and (per_rank > 0).sum() == 1


def test_merge_results_joins_spanning_nodes_in_rank_order():
    from schwarzwald_b200 import distributed
    from schwarzwald_b200.tiler import NODE_DTYPE, TileResult
    a = TileResult(np.array([(0, 0, 0, 0, 2), (5, 1, 0, 2, 3)], NODE_DTYPE), np.array([10, 11, 1, 2, 3], np.uint32), 1)
    b = TileResult(np.array([(6, 1, 1, 0, 1), (0, 0, 0, 1, 2)], NODE_DTYPE), np.array([7, 20, 21], np.uint32), 1)
    m = distributed.merge_results([a, b])
    d = m.as_dict()
    assert list(d["r"]) == [10, 11, 20, 21]
    assert list(d["r5"]) == [1, 2, 3] and list(d["r6"]) == [7]
    assert m.start_level == 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from schwarzwald_b200 import distributed
dist.init_process_group("gloo")
comm = distributed.TorchDistComm()
rank, world = comm.rank, comm.world
rng = np.random.default_rng(100 + rank)
keys = rng.integers(0, 2 ** 63, 20000, dtype=np.uint64)
bins = torch.from_numpy(np.bincount((keys >> np.uint64(45)).astype(np.int64), minlength=262144).astype(np.int32))
comm.all_reduce_sum(bins)
assert int(bins.sum()) == 20000 * world
fp = distributed.choose_splitters(bins.numpy().view(np.uint32), world, 6)
dest = np.searchsorted(fp[1:-1], (keys >> np.uint64(45)).astype(np.uint32), side="right")
order = np.argsort(dest, kind="stable")
send = torch.from_numpy(keys[order].view(np.int64).copy()).reshape(-1, 1)
counts = np.bincount(dest, minlength=world)
recv, recv_counts = comm.all_to_all_rows(send, counts.tolist())
got = recv.numpy().view(np.uint64).ravel()
pref = (got >> np.uint64(45)).astype(np.uint32)
assert ((pref >= fp[rank]) & (pref < fp[rank + 1])).all(), "received a point of a foreign subtree"
tot = torch.tensor([len(got)]); comm.all_reduce_sum(tot)
assert int(tot) == 20000 * world
# sources arrive in rank order: the own part sits at offset sum(recv_counts[:rank])
off = sum(recv_counts[:rank])
mine = keys[order][dest[order] == rank]
assert np.array_equal(got[off:off + recv_counts[rank]], mine)
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_histogram_splitters_and_exchange(tmp_path):
    """world_size 2 over gloo: summed histogram, splitters, all-to-all row exchange."""
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_thread_comm_matches_expected_semantics():
    import threading
    import torch
    from schwarzwald_b200 import distributed
    comms = distributed.ThreadComm.create(3)
    out = [None] * 3

    def work(r):
        t = torch.full((4,), r + 1, dtype=torch.int32)
        comms[r].all_reduce_sum(t)
        send = torch.arange(6, dtype=torch.int64).reshape(6, 1) + 100 * r
        recv, rc = comms[r].all_to_all_rows(send, [1, 2, 3])
        out[r] = (t.clone(), recv.ravel().tolist(), rc)

    th = [threading.Thread(target=work, args=(r,)) for r in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(int(o[0][0]) == 6 for o in out)
    assert out[0][1] == [0, 100, 200] and out[0][2] == [1, 1, 1]
    assert out[1][1] == [1, 2, 101, 102, 201, 202]
    assert out[2][1] == [3, 4, 5, 103, 104, 105, 203, 204, 205]
