"""N > 1 host logic on CPU: two gloo ranks fold their shard of a result set into the dense per-taxon vector
and all-reduce it; the sum must equal the fold of the whole set (what one rank would have reported)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centrifuge_b200.abundance import taxon_counts

REC = np.dtype([("taxid", "<u8"), ("score", "<u4"), ("hitlen", "<u4"), ("uid", "<u4"), ("pad", "<u4")])


def make_results(n_units, seed):
    rng = np.random.default_rng(seed)
    cnt = rng.choice([0, 1, 1, 1, 2, 3, 5], size=n_units)
    off = np.zeros(n_units + 1, dtype=np.uint32)
    off[1:] = np.cumsum(cnt)
    recs = np.zeros(int(off[-1]), dtype=REC)
    recs["taxid"] = rng.choice(np.array([1000, 1001, 1002, 1003, 100, 101, 77777], dtype=np.uint64), size=len(recs))
    recs["score"] = rng.choice([400, 7225, 7225, 2441], size=len(recs))
    return off, recs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes = np.array([1, 100, 101, 1000, 1001, 1002, 1003], dtype=np.uint64)
    off, recs = make_results(5000, 100 + rank)              # this rank's shard of the read stream
    t = torch.from_numpy(taxon_counts(nodes, off, recs, k=5))
    dist.all_reduce(t)                                      # the one collective of the path
    if rank == 0:
        q.put(t.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_taxon_count_allreduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nodes = np.array([1, 100, 101, 1000, 1001, 1002, 1003], dtype=np.uint64)
    exp = sum(taxon_counts(nodes, *make_results(5000, 100 + r), k=5) for r in range(2))
    assert np.array_equal(got, exp)


def test_taxon_counts_matches_reference_rule():
    """Against a literal per-unit loop of the reference's rule (top-score streak, capped at k)."""
    nodes = np.array([1, 100, 101, 1000, 1001, 1002, 1003], dtype=np.uint64)
    off, recs = make_results(3000, 7)
    for k in (1, 2, 5):
        exp = np.zeros((len(nodes) + 1, 2), dtype=np.int64)
        for u in range(len(off) - 1):
            r = recs[off[u]:off[u + 1]]
            if len(r) == 0:
                exp[-1] += 1
                continue
            best = r["score"].max()
            rep = r[r["score"] == best][:k]
            for t in rep["taxid"]:
                i = np.searchsorted(nodes, t)
                i = i if i < len(nodes) and nodes[i] == t else len(nodes)
                exp[i, 0] += 1
                if len(rep) == 1:
                    exp[i, 1] += 1
        assert np.array_equal(taxon_counts(nodes, off, recs, k=k), exp)
