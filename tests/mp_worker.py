"""Worker launched by the multi-process tests: `python -m torch.distributed.run ... tests/mp_worker.py MODE`.

MODE cpu : gloo, world_size 2 -- the host-side exchange logic of dblink_b200.distributed on CPU tensors.
MODE gpu : nccl, one rank per GPU -- the sharded chain must equal the oracle chain bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cpu_mode():
    import torch
    import torch.distributed as dist

    from dblink_b200.distributed import allreduce_summary, exchange, gather_blobs, lpt_assign, merge_owned, sum_hashes

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cpu")
    A = 3
    ew = A + 1
    # rank r sends (r + d + 1) entity messages and 2*(r+d) record messages to rank d; payload encodes (src, dst, i)
    ec = np.array([rank + d + 1 for d in range(world)], np.int64)
    rc = np.array([2 * (rank + d) for d in range(world)], np.int64)
    ent = [[rank * 1000 + d * 100 + i] + [rank, d, i] for d in range(world) for i in range(ec[d])]
    rec = [[rank * 1000 + d * 100 + i, d, rank] for d in range(world) for i in range(rc[d])]
    send_ent = torch.tensor(np.array(ent, np.int32).reshape(-1), dtype=torch.int32)
    send_rec = torch.tensor(np.array(rec, np.int32).reshape(-1) if rec else np.zeros(0, np.int32), dtype=torch.int32)
    recv_ent, ne, recv_rec, nr = exchange(dist, send_ent, ec, send_rec, rc, ew, dev, torch)
    assert ne == sum(s + rank + 1 for s in range(world)) and nr == sum(2 * (s + rank) for s in range(world))
    got = recv_ent.numpy().reshape(-1, ew)
    exp = [[s * 1000 + rank * 100 + i, s, rank, i] for s in range(world) for i in range(s + rank + 1)]
    assert got.tolist() == exp, (got.tolist(), exp)
    gotr = recv_rec.numpy().reshape(-1, 3)
    expr = [[s * 1000 + rank * 100 + i, rank, s] for s in range(world) for i in range(2 * (s + rank))]
    assert gotr.tolist() == expr
    counts, ll, failed = allreduce_summary(dist, np.arange(5, dtype=np.int64) * (rank + 1), 1.5 * (rank + 1), dev, torch,
                                           failed=int(rank == 1))
    tot = sum(r + 1 for r in range(world))
    assert counts.tolist() == (np.arange(5) * tot).tolist() and abs(ll - 1.5 * tot) < 1e-12
    assert failed == 1  # an error on one rank is seen by every rank after the collective
    # the handshake of the peer-to-peer data plane: fixed-size descriptions, gathered in rank order
    blob = bytes([rank] * 192)
    blobs = gather_blobs(dist, blob, dev, torch)
    assert len(blobs) == 192 * world and all(blobs[192 * r:192 * (r + 1)] == bytes([r] * 192) for r in range(world))
    # row fingerprints add up mod 2^64 on every rank
    big = (1 << 64) - 5
    he, hr = sum_hashes(dist, big if rank == 0 else 7, 11 * (rank + 1), dev, torch)
    exp_e = (big + 7 * (world - 1)) % (1 << 64)
    assert he == exp_e and hr == 11 * sum(r + 1 for r in range(world))
    # owned rows of every rank -> the full state
    R, E, A = 10, 8, 3
    full_y = np.arange(E * A, dtype=np.int32).reshape(E, A)
    full_l = (np.arange(R, dtype=np.int32) * 3) % E
    full_z = (np.arange(R * A).reshape(R, A) % 2).astype(np.uint8)
    parts = []
    for r in range(world):
        e = np.arange(r, E, world)
        rr = np.arange(r, R, world)
        parts.append({"ent_ids": e, "y": full_y[e], "block": e % 4, "rec_ids": rr, "link": full_l[rr], "z": full_z[rr]})
    m = merge_owned(parts, R, E, A)
    assert np.array_equal(m["y"], full_y) and np.array_equal(m["link"], full_l) and np.array_equal(m["z"], full_z)
    assert np.array_equal(m["block"], np.arange(E) % 4)
    owner = lpt_assign([9, 1, 8, 2, 7, 3, 3, 3], world)
    loads = np.bincount(owner, weights=[9, 1, 8, 2, 7, 3, 3, 3], minlength=world)
    assert loads.max() - loads.min() <= 3 and set(owner.tolist()) == set(range(world))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("cpu exchange ok")


def gpu_mode():
    import torch
    import torch.distributed as dist

    from helpers import oracle_setup, synth_problem
    from oracle import oracle as O
    from dblink_b200 import synth
    from dblink_b200.distributed import ShardedGibbs

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = synth_problem(seed=5, R=1500, n_files=2)
    import dblink_b200 as D

    rc = D.RecordsCache.build(g["values"], g["files"], g["attributes"])
    x, file = rc.transform_records(g["values"], g["files"])
    alpha = [a.alpha for a in g["attributes"]]
    beta = [a.beta for a in g["attributes"]]
    # the peer-to-peer data plane (CUDA IPC between the rank processes) for every sampler, and the host-mediated
    # fallback (NCCL all-to-alls driven from here) once
    for sampler, mode in (("PCG-II", "p2p"), ("PCG-I", "p2p"), ("Gibbs", "p2p"), ("PCG-II", "host")):
        eng = ShardedGibbs(rc.indexes, alpha, beta, seed=99, num_files=len(rc.file_ids), levels=3, split_attrs=(2, 3),
                           exchange=mode)
        eng.init_state(x, file)
        if mode == "p2p" and world > 1:
            assert eng.connected, "peer mapping failed: the sharded sweep would fall back to the host path"
        m, st, tree, ox, ofile = oracle_setup(O, g, 99, 3, (2, 3))
        moved = 0
        for it in range(5):
            eng.sweep(sampler, 1)
            moved += eng.last_exchange[0]
            st.sweep(O.SAMPLERS[sampler])
            d = eng.download_state()
            for k in ("theta", "link", "y", "z"):
                assert np.array_equal(d[k], getattr(st, k)), (sampler, it, k, rank)
            ps, os_ = eng.summary(), st.summary()
            assert ps["num_isolates"] == os_["num_isolates"]
            assert np.array_equal(ps["agg_dist"], os_["agg_dist"]) and np.array_equal(ps["rec_dist"], os_["rec_dist"])
            assert abs(ps["log_likelihood"] - os_["log_likelihood"]) <= 1e-9 * abs(os_["log_likelihood"])
        # resume from host arrays (sliced upload + all-gather) into pinned output buffers, then keep following the oracle
        pinned = {k: torch.empty(d[k].shape, dtype=torch.from_numpy(d[k][:0].copy()).dtype, pin_memory=True).numpy()
                  for k in ("z", "link", "y", "block")}
        eng.upload_state(x, file, d["z"], d["link"], d["y"], d["theta"], iteration=eng.iteration)
        for it in range(2):
            eng.sweep(sampler, 1)
            st.sweep(O.SAMPLERS[sampler])
            d = eng.download_state(out=pinned)
            for k in ("theta", "link", "y", "z"):
                assert np.array_equal(d[k], getattr(st, k)), ("after upload", sampler, it, k, rank)
        t = torch.tensor([moved], device="cuda")
        dist.all_reduce(t)
        if world > 1:
            assert int(t[0]) > 0, "the test should move clusters between ranks"
        # rank-count-invariant fingerprint == the one of the oracle's state
        from helpers import state_hash_numpy
        from dblink_b200.engine import combine_state_hash

        assert eng.state_hash() == combine_state_hash(*state_hash_numpy(st.y, st.link, st.z), st.theta, st.iteration)
        eng.eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"gpu sharded chain == oracle chain on {world} ranks")


def sample_mode():
    """Sampler.sample end to end on a sharded engine: every rank sweeps, rank 0 writes the outputs."""
    import tempfile

    import torch
    import torch.distributed as dist

    from helpers import synth_problem
    import dblink_b200 as D
    from dblink_b200 import sampler as chain
    from dblink_b200.distributed import ShardedGibbs

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = synth_problem(seed=6, R=900, n_files=2)
    rc = D.RecordsCache.build(g["values"], g["files"], g["attributes"])
    x, file = rc.transform_records(g["values"], g["files"])
    eng = ShardedGibbs(rc.indexes, [a.alpha for a in g["attributes"]], [a.beta for a in g["attributes"]], seed=3,
                       num_files=len(rc.file_ids), levels=2, split_attrs=(2, 3))
    eng.init_state(x, file)
    out = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(out, src=0)
    n = chain.sample(eng, [f"r{i}" for i in range(len(x))], [a.name for a in g["attributes"]], 5, out[0],
                     burnin_interval=2, thinning_interval=3, sampler="PCG-I")
    assert n == 2 + 4 * 3 and eng.iteration == n
    dist.barrier()
    if rank == 0:
        import pyarrow.parquet as pq

        t = pq.read_table(os.path.join(out[0], "linkage-chain.parquet")).to_pandas()
        assert sorted(set(t["iteration"])) == [2, 5, 8, 11, 14]
        rows = open(os.path.join(out[0], "diagnostics.csv")).read().strip().splitlines()
        assert len(rows) == 1 + 5
        print(f"sharded sample ok on {world} ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    {"cpu": cpu_mode, "gpu": gpu_mode, "sample": sample_mode}[sys.argv[1]]()
