"""Posterior summaries of a linkage chain (host side; small inputs).

  most_probable_clusters / shared_most_probable_clusters  <- LinkageChain.scala:52-109
  cluster_size_distribution / partition_sizes             <- LinkageChain.scala:118-154
  pairwise_metrics                                        <- analysis/PairwiseMetrics.scala:44-63,
                                                             BinaryClassificationMetrics.scala:23-37
  adjusted_rand_index                                     <- analysis/ClusteringMetrics.scala:41-74

A linkage chain is a list of samples; a sample is (iteration, {partition_id: [cluster, ...]}) with clusters as
collections of record ids (what linkage-chain.parquet stores, package.scala:94-96).
"""
import collections
import itertools
import math


def clusters_of_sample(sample):
    for clusters in sample[1].values():
        for c in clusters:
            if len(c):
                yield frozenset(c)


def most_probable_clusters(chain):
    """record id -> (cluster, frequency): the cluster the record appears in most often along the chain
    (LinkageChain.scala:52-64; ties are broken by the reduce order there, by first occurrence here)."""
    n = len({s[0] for s in chain})
    freq = collections.Counter()
    order = {}
    for s in chain:
        for c in clusters_of_sample(s):
            freq[c] += 1.0 / n
            order.setdefault(c, len(order))
    best = {}
    for c, f in sorted(freq.items(), key=lambda kv: order[kv[0]]):
        for r in c:
            if r not in best or f > best[r][1]:
                best[r] = (c, f)
    return best


def shared_most_probable_clusters(chain_or_mpc):
    """Records grouped by their most probable cluster (LinkageChain.scala:75-95; the reference's stricter
    'shared' filter is commented out there, :88-94, and is not applied here either)."""
    mpc = chain_or_mpc if isinstance(chain_or_mpc, dict) else most_probable_clusters(chain_or_mpc)
    groups = collections.defaultdict(set)
    for r, (c, _) in mpc.items():
        groups[c].add(r)
    return [frozenset(v) for v in groups.values()]


def cluster_size_distribution(chain):
    """iteration -> {cluster size: count} (LinkageChain.scala:137-154)."""
    out = {}
    for it, parts in chain:
        d = collections.Counter()
        for clusters in parts.values():
            for c in clusters:
                d[len(c)] += 1
        out[it] = dict(d)
    return out


def partition_sizes(chain):
    """iteration -> {partition id: number of clusters} (LinkageChain.scala:118-128)."""
    return {it: {p: len(cl) for p, cl in parts.items()} for it, parts in chain}


def to_pairwise_links(clusters):
    links = set()
    for c in clusters:
        for a, b in itertools.combinations(sorted(c), 2):
            links.add((a, b))
    return links


def pairwise_metrics(predicted_clusters, true_clusters):
    """precision / recall / F1 over record pairs (PairwiseMetrics.scala:44-63)."""
    pred, true = to_pairwise_links(predicted_clusters), to_pairwise_links(true_clusters)
    tp = len(pred & true)
    fp, fn = len(pred) - tp, len(true) - tp
    precision = tp / (tp + fp) if tp + fp else float("nan")
    recall = tp / (tp + fn) if tp + fn else float("nan")
    f1 = 2 * precision * recall / (precision + recall) if tp else (0.0 if (fp or fn) else float("nan"))
    return {"precision": precision, "recall": recall, "f1score": f1, "TP": tp, "FP": fp, "FN": fn}


def adjusted_rand_index(predicted_clusters, true_clusters):
    """ClusteringMetrics.AdjustedRandIndex (ClusteringMetrics.scala:44-74)."""
    comb2 = lambda x: x * (x - 1) // 2 if x >= 2 else 0  # noqa: E731
    pred_of, true_of = {}, {}
    for i, c in enumerate(predicted_clusters):
        for r in c:
            pred_of[r] = i
    for i, c in enumerate(true_clusters):
        for r in c:
            true_of[r] = i
    if set(pred_of) != set(true_of):
        raise ValueError("predicted and true clusterings must cover the same records")
    table = collections.Counter((pred_of[r], true_of[r]) for r in pred_of)
    pred_sum, true_sum = collections.Counter(), collections.Counter()
    for (p, t), n in table.items():
        pred_sum[p] += n
        true_sum[t] += n
    pc = sum(comb2(v) for v in pred_sum.values())
    tc = sum(comb2(v) for v in true_sum.values())
    total = sum(comb2(v) for v in table.values())
    expected = pc * tc / comb2(len(pred_of))
    max_index = (pc + tc) / 2.0
    return (total - expected) / (max_index - expected) if max_index != expected else 1.0


def membership_to_clusters(record_ids, membership):
    groups = collections.defaultdict(set)
    for r, m in zip(record_ids, membership):
        groups[m].add(r)
    return [frozenset(v) for v in groups.values()]


def format_pairwise(m):
    return ("=====================================\n        Pairwise metrics              \n"
            "-------------------------------------\n"
            f" Precision:      {m['precision']}\n Recall:         {m['recall']}\n F1-score:       {m['f1score']}\n"
            "=====================================\n")


def format_cluster(ari):
    return ("=====================================\n          Cluster metrics            \n"
            "-------------------------------------\n"
            f" Adj. Rand index: {ari}\n=====================================\n")


def is_nan(x):
    return isinstance(x, float) and math.isnan(x)
