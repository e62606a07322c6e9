"""Chain outputs in the reference's formats.

  LinkageChainWriter  linkage-chain.parquet, hive-partitioned by partitionId, columns iteration:int64,
                      linkageStructure:list<list<string>> (package.scala:94-96, util/BufferedRDDWriter.scala:44-50)
  DiagnosticsWriter   diagnostics.csv with the header of DiagnosticsWriter.scala:39-45 and rows of :47-72
  save_cluster_size_distribution / save_partition_sizes   LinkageChain.scala:162-211
"""
import os
import time

import numpy as np


def linkage_structure(link, block_of_entity, record_ids):
    """State.getLinkageStructure (State.scala:102-112): partition id -> list of clusters (lists of record ids);
    isolated entities carry no cluster."""
    order = np.argsort(link, kind="stable")
    sl = link[order]
    starts = np.flatnonzero(np.r_[True, sl[1:] != sl[:-1]])
    ends = np.r_[starts[1:], len(sl)]
    parts = {}
    for s, e in zip(starts, ends):
        ent = sl[s]
        parts.setdefault(int(block_of_entity[ent]), []).append([record_ids[i] for i in order[s:e]])
    return parts


def linkage_structure_arrow(link, block_of_entity, record_ids):
    """The same structure without Python loops over clusters: partition id -> pyarrow ListArray (list<string>), one
    entry per non-isolated entity, clusters in ascending entity id, records of a cluster in ascending record index.
    `record_ids` = a pyarrow string array (or anything pa.array accepts).  At 1M records this takes ~0.2 s where
    the list-of-lists form takes seconds, so a recorded sample costs about as much as the sweeps between samples."""
    import pyarrow as pa

    link = np.asarray(link)
    R = len(link)
    ids = record_ids if isinstance(record_ids, pa.Array) else pa.array([str(r) for r in record_ids], pa.string())
    rec_blk = np.asarray(block_of_entity)[link]
    order = np.lexsort((link, rec_blk))  # stable: by block, then entity, then record index
    sl, sb = link[order], rec_blk[order]
    cstart = np.flatnonzero(np.r_[True, sl[1:] != sl[:-1]]) if R else np.zeros(0, np.int64)
    cend = np.r_[cstart[1:], R]
    cblk = sb[cstart]
    values = ids.take(pa.array(order, pa.int64()))
    blocks, first = np.unique(cblk, return_index=True)
    last = np.r_[first[1:], len(cstart)]
    parts = {}
    for b, f, e in zip(blocks, first, last):  # one iteration per partition
        lo, hi = int(cstart[f]), int(cend[e - 1])
        offs = np.r_[cstart[f:e], hi] - lo
        parts[int(b)] = pa.ListArray.from_arrays(pa.array(offs, pa.int32()), values.slice(lo, hi - lo))
    # a partition that holds only isolated entities still has a row, with an empty linkage structure
    # (State.getLinkageStructure maps over every partition, State.scala:102-112)
    for b in np.unique(np.asarray(block_of_entity)):
        if int(b) not in parts:
            parts[int(b)] = pa.ListArray.from_arrays(pa.array([0], pa.int32()), pa.array([], pa.string()))
    return parts


class LinkageChainWriter:
    def __init__(self, path, write_buffer_size=10, append=False):
        import pyarrow as pa

        self.pa = pa
        self.path = path
        self.buf = []
        self.n = write_buffer_size
        self.file_no = 0
        self.schema = pa.schema([("iteration", pa.int64()), ("linkageStructure", pa.list_(pa.list_(pa.string())))])
        if os.path.exists(path) and not append:
            import shutil

            shutil.rmtree(path)
        os.makedirs(path, exist_ok=True)
        if append:
            self.file_no = sum(len(f) for _, _, f in os.walk(path))

    def append(self, iteration, parts):
        self.buf.append((iteration, parts))
        if len(self.buf) >= self.n:
            self.flush()

    def flush(self):
        import pyarrow.parquet as pq

        if not self.buf:
            return
        by_part = {}
        for it, parts in self.buf:
            for pid, clusters in parts.items():
                by_part.setdefault(pid, []).append((it, clusters))
        for pid, rows in by_part.items():
            d = os.path.join(self.path, f"partitionId={pid}")
            os.makedirs(d, exist_ok=True)
            pa = self.pa
            its = pa.array([r[0] for r in rows], pa.int64())
            if all(isinstance(r[1], pa.Array) for r in rows):  # linkage_structure_arrow: no Python lists at all
                offs = np.r_[0, np.cumsum([len(r[1]) for r in rows])]
                ls = pa.ListArray.from_arrays(pa.array(offs, pa.int32()), pa.concat_arrays([r[1] for r in rows]))
            else:
                ls = pa.array([r[1].to_pylist() if isinstance(r[1], pa.Array) else r[1] for r in rows],
                              pa.list_(pa.list_(pa.string())))
            tbl = pa.Table.from_arrays([its, ls], schema=self.schema)
            pq.write_table(tbl, os.path.join(d, f"part-{self.file_no:05d}.parquet"))
        self.file_no += 1
        self.buf = []

    close = flush


def read_linkage_chain(path, lower_iteration_cutoff=0):
    """LinkageChain.readLinkageChain (LinkageChain.scala:35-43) -> [(iteration, {partitionId: clusters})] sorted."""
    import pyarrow.parquet as pq

    samples = {}
    for d in sorted(os.listdir(path)):
        if not d.startswith("partitionId="):
            continue
        pid = int(d.split("=")[1])
        for f in sorted(os.listdir(os.path.join(path, d))):
            t = pq.ParquetFile(os.path.join(path, d, f)).read().to_pydict()
            for it, ls in zip(t["iteration"], t["linkageStructure"]):
                if it >= lower_iteration_cutoff:
                    samples.setdefault(it, {})[pid] = ls
    return sorted(samples.items())


class DiagnosticsWriter:
    def __init__(self, path, attribute_names, append=False):
        self.path = path
        self.names = list(attribute_names)
        new = not (append and os.path.exists(path))
        self.fh = open(path, "a" if append else "w")
        if new:
            agg = ",".join(f"aggDist-{n}" for n in self.names)
            rec = ",".join(f"recDistortion-{k}" for k in range(len(self.names) + 1))
            self.fh.write(f"iteration,systemTime-ms,numObservedEntities,logLikelihood,popSize,{agg},{rec}\n")

    def write_row(self, summary, pop_size):
        agg = summary["agg_dist"].sum(axis=1)  # summed over files (DiagnosticsWriter.scala:52-54)
        row = [str(summary["iteration"]), str(int(time.time() * 1000)), str(pop_size - summary["num_isolates"]),
               f"{summary['log_likelihood']:.9e}", str(pop_size)]
        row += [str(int(v)) for v in agg] + [str(int(v)) for v in summary["rec_dist"]]
        self.fh.write(",".join(row) + "\n")

    def close(self):
        self.fh.close()


def save_cluster_size_distribution(dist, path):
    """LinkageChain.saveClusterSizeDistribution (LinkageChain.scala:162-185)."""
    its = sorted(dist)
    mx = max((max(d) if d else 0) for d in dist.values()) if dist else 0
    with open(os.path.join(path, "cluster-size-distribution.csv"), "w") as fh:
        fh.write("iteration," + ",".join(str(k) for k in range(mx + 1)) + "\n")
        for it in its:
            fh.write(str(it) + "," + ",".join(str(dist[it].get(k, 0)) for k in range(mx + 1)) + "\n")


def save_partition_sizes(sizes, path):
    """LinkageChain.savePartitionSizes (LinkageChain.scala:193-211)."""
    its = sorted(sizes)
    pids = sorted({p for d in sizes.values() for p in d})
    with open(os.path.join(path, "partition-sizes.csv"), "w") as fh:
        fh.write("iteration," + ",".join(str(p) for p in pids) + "\n")
        for it in its:
            fh.write(str(it) + "," + ",".join(str(sizes[it].get(p, 0)) for p in pids) + "\n")
