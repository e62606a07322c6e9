"""Posterior summaries on arrays (no Python loop over clusters or records): the same quantities as analysis.py --
LinkageChain.scala:52-154 and analysis/{PairwiseMetrics,ClusteringMetrics}.scala -- for chains of millions of records.

A sample is held as (members, offsets, partition_of_cluster): `members` = record INDICES (positions in the id
dictionary) of all clusters back to back, `offsets[c]:offsets[c+1]` = cluster c.  The simple set-based functions in
analysis.py stay as the readable restatement; tests compare the two on random chains.
"""
import os

import numpy as np


# ---- reading the chain -------------------------------------------------------------------------------
class ChainArrays:
    def __init__(self, record_ids, iterations, samples):
        self.record_ids = record_ids  # pyarrow string array: index -> record id
        self.iterations = iterations  # int64[S], ascending
        self.samples = samples        # list of (members int32[n], offsets int64[nc+1], partition int32[nc])

    @property
    def num_records(self):
        return len(self.record_ids)


def read_chain_arrays(path, lower_iteration_cutoff=0):
    """linkage-chain.parquet (hive-partitioned by partitionId) -> ChainArrays, through Arrow only."""
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.parquet as pq

    rows = {}  # iteration -> list of (pid, ListArray of clusters)
    for d in sorted(os.listdir(path)):
        if not d.startswith("partitionId="):
            continue
        pid = int(d.split("=")[1])
        for f in sorted(os.listdir(os.path.join(path, d))):
            t = pq.ParquetFile(os.path.join(path, d, f)).read()
            its = t.column("iteration").to_numpy()
            ls = t.column("linkageStructure").combine_chunks()
            for i, it in enumerate(its):
                if it >= lower_iteration_cutoff:
                    rows.setdefault(int(it), []).append((pid, ls[i].values))  # list<string> of that row
    its = sorted(rows)
    if not its:
        return ChainArrays(pa.array([], pa.string()), np.zeros(0, np.int64), [])
    first = pa.concat_arrays([cl.flatten() for _, cl in rows[its[0]]])
    ids = pc.unique(first)
    samples = []
    for it in its:
        mem, sizes, part = [], [], []
        for pid, cl in rows[it]:
            flat = cl.flatten()
            idx = pc.index_in(flat, value_set=ids)
            if idx.null_count:
                raise ValueError("a sample mentions a record id that the first sample does not")
            mem.append(idx.to_numpy(zero_copy_only=False).astype(np.int32))
            off = cl.offsets.to_numpy().astype(np.int64)
            sz = np.diff(off)
            sizes.append(sz)
            part.append(np.full(len(sz), pid, np.int32))
        sizes = np.concatenate(sizes) if sizes else np.zeros(0, np.int64)
        keep = sizes > 0
        samples.append((np.concatenate(mem) if mem else np.zeros(0, np.int32),
                        np.r_[0, np.cumsum(sizes[keep])].astype(np.int64),
                        (np.concatenate(part) if part else np.zeros(0, np.int32))[keep]))
    return ChainArrays(ids, np.asarray(its, np.int64), samples)


def sample_from_links(link, block_of_entity):
    """One sample straight from the engine's arrays (record index = position): (members, offsets, partition)."""
    link = np.asarray(link)
    order = np.argsort(link, kind="stable").astype(np.int32)
    sl = link[order]
    start = np.flatnonzero(np.r_[True, sl[1:] != sl[:-1]]) if len(sl) else np.zeros(0, np.int64)
    return order, np.r_[start, len(sl)].astype(np.int64), np.asarray(block_of_entity)[sl[start]].astype(np.int32)


# ---- LinkageChain.scala:118-154 ----------------------------------------------------------------------
def cluster_size_distribution(chain):
    """iteration -> {cluster size: count}."""
    out = {}
    for it, (_, off, _) in zip(chain.iterations, chain.samples):
        sz = np.diff(off)
        cnt = np.bincount(sz)
        out[int(it)] = {int(k): int(cnt[k]) for k in np.flatnonzero(cnt)}
    return out


def partition_sizes(chain):
    """iteration -> {partition id: number of clusters}."""
    out = {}
    for it, (_, _, part) in zip(chain.iterations, chain.samples):
        p, n = np.unique(part, return_counts=True)
        out[int(it)] = {int(a): int(b) for a, b in zip(p, n)}
    return out


# ---- most probable clusters (LinkageChain.scala:52-109) ------------------------------------------------
def _mix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def cluster_signatures(chain):
    """uint64[S, R]: for every sample and record, a 64-bit signature of the SET of records it is clustered with
    (sum of mixed member indices, mixed with the size); equal sets <=> equal signatures up to 2^-64 collisions."""
    R = chain.num_records
    mix = _mix64(np.arange(R, dtype=np.uint64))
    sig = np.zeros((len(chain.samples), R), np.uint64)
    for s, (mem, off, _) in enumerate(chain.samples):
        if len(mem) != R:
            raise ValueError("every sample must mention every record exactly once")
        sizes = np.diff(off)
        with np.errstate(over="ignore"):
            h = np.add.reduceat(mix[mem], off[:-1]) if len(sizes) else np.zeros(0, np.uint64)
            h = _mix64(h ^ _mix64(sizes.astype(np.uint64)))
        sig[s, mem] = np.repeat(h, sizes)
    return sig


def most_probable_signature(sig, block=65536):
    """For every record the signature it carries most often along the chain (ties: the one seen first) and its
    frequency: (uint64[R], float64[R]).  Works on blocks of records to bound memory."""
    S, R = sig.shape
    best = np.zeros(R, np.uint64)
    freq = np.zeros(R)
    for lo in range(0, R, block):
        a = sig[:, lo:lo + block]
        n = a.shape[1]
        order = np.argsort(a, axis=0, kind="stable")       # equal signatures keep ascending sample order
        srt = np.take_along_axis(a, order, 0).T.reshape(-1)  # record-major: the S signatures of a record are adjacent
        first_sample = order.T.reshape(-1)
        new = np.ones(n * S, bool)
        new[1:] = srt[1:] != srt[:-1]
        new[::S] = True                                      # a run never crosses records
        run_start = np.flatnonzero(new)
        count = np.diff(np.r_[run_start, n * S])
        # winner per record: larger count first, then the earlier first sample (a lexicographic sort of the runs; a
        # packed integer key would overflow int64 for chains of ~50k samples)
        rec_of_run = run_start // S
        by = np.lexsort((first_sample[run_start], -count, rec_of_run))
        lead = np.ones(len(by), bool)
        lead[1:] = rec_of_run[by][1:] != rec_of_run[by][:-1]
        win = by[lead]                                       # one run per record, in record order
        best[lo:lo + n] = srt[run_start[win]]
        freq[lo:lo + n] = count[win] / S
    return best, freq


def shared_most_probable_clusters(chain):
    """Records grouped by their most probable cluster: int32 labels[R] (records with equal labels form a cluster),
    labelled by the smallest record index of the group."""
    sig = cluster_signatures(chain)
    best, _ = most_probable_signature(sig)
    _, inv = np.unique(best, return_inverse=True)
    rep = np.full(inv.max() + 1 if len(inv) else 0, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(rep, inv, np.arange(len(inv)))
    return rep[inv].astype(np.int64)


def labels_to_clusters(labels, record_ids=None):
    """labels -> list of clusters (arrays of record indices, or lists of ids when record_ids is given)."""
    order = np.argsort(labels, kind="stable")
    sl = np.asarray(labels)[order]
    start = np.flatnonzero(np.r_[True, sl[1:] != sl[:-1]]) if len(sl) else np.zeros(0, np.int64)
    groups = np.split(order, start[1:])
    if record_ids is None:
        return groups
    ids = np.asarray(record_ids.to_pylist() if hasattr(record_ids, "to_pylist") else record_ids, dtype=object)
    return [list(ids[g]) for g in groups]


# ---- metrics (PairwiseMetrics.scala:44-63, ClusteringMetrics.scala:44-74) ------------------------------
def _comb2(x):
    x = np.asarray(x, np.int64)
    return x * (x - 1) // 2


def contingency(pred_labels, true_labels):
    _, p = np.unique(pred_labels, return_inverse=True)
    _, t = np.unique(true_labels, return_inverse=True)
    nt = int(t.max()) + 1 if len(t) else 0
    _, n = np.unique(p.astype(np.int64) * max(nt, 1) + t, return_counts=True)
    return n, np.bincount(p), np.bincount(t)


def pairwise_metrics(pred_labels, true_labels):
    n, pn, tn = contingency(pred_labels, true_labels)
    tp = int(_comb2(n).sum())
    fp = int(_comb2(pn).sum()) - tp
    fn = int(_comb2(tn).sum()) - tp
    precision = tp / (tp + fp) if tp + fp else float("nan")
    recall = tp / (tp + fn) if tp + fn else float("nan")
    f1 = 2 * precision * recall / (precision + recall) if tp else (0.0 if (fp or fn) else float("nan"))
    return {"precision": precision, "recall": recall, "f1score": f1, "TP": tp, "FP": fp, "FN": fn}


def adjusted_rand_index(pred_labels, true_labels):
    n, pn, tn = contingency(pred_labels, true_labels)
    total, pc, tc = int(_comb2(n).sum()), int(_comb2(pn).sum()), int(_comb2(tn).sum())
    expected = pc * tc / int(_comb2(len(pred_labels)))
    max_index = (pc + tc) / 2.0
    return (total - expected) / (max_index - expected) if max_index != expected else 1.0
