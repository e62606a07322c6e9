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
            tbl = self.pa.table({"iteration": [r[0] for r in rows], "linkageStructure": [r[1] for r in rows]},
                                schema=self.schema)
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
