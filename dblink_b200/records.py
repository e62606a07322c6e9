"""RecordsCache: statistics + attribute indexes for a collection of records (RecordsCache.scala:31-134).

Host-side mirror: value counts per attribute in one pass (RecordsCache.scala:75-96), one AttributeIndex per
matching attribute (:104-114), records transformed to integer value ids with -1 for missing (:117-133).
"""
import collections
import csv
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .engine import AttributeIndex


@dataclass
class SimilarityFn:
    """SimilarityFn.scala:50-107."""
    name: str = "ConstantSimilarityFn"
    threshold: float = 7.0
    max_similarity: float = 10.0

    def __post_init__(self):
        if self.name not in ("ConstantSimilarityFn", "LevenshteinSimilarityFn"):
            raise ValueError(f"unsupported similarity function {self.name}")
        if self.name != "ConstantSimilarityFn":
            if not self.max_similarity > 0.0:
                raise ValueError("`maxSimilarity` must be positive")  # SimilarityFn.scala:59
            if not (0.0 <= self.threshold < self.max_similarity):
                raise ValueError("`threshold` must be in the interval [0, maxSimilarity)")  # :60-61

    @property
    def is_constant(self):
        return self.name == "ConstantSimilarityFn"


@dataclass
class Attribute:
    """package.scala:122-133 + BetaShapeParameters :164-168."""
    name: str
    similarity_fn: SimilarityFn = field(default_factory=SimilarityFn)
    alpha: float = 1.0
    beta: float = 1.0

    def __post_init__(self):
        if not (self.alpha > 0 and self.beta > 0):
            raise ValueError("shape parameters must be positive")

    @property
    def is_constant(self):
        return self.similarity_fn.is_constant


class RecordsCache:
    def __init__(self, attributes: List[Attribute], indexes: List[AttributeIndex], file_ids: List[str],
                 file_sizes: List[int], missing_counts=None):
        self.attributes = attributes
        self.indexes = indexes
        self.file_ids = file_ids          # sorted file id strings; position = integer file id
        self.file_sizes = file_sizes
        self.missing_counts = missing_counts

    @property
    def num_attributes(self):
        return len(self.attributes)

    @property
    def num_records(self):
        return int(sum(self.file_sizes))

    @classmethod
    def build(cls, values: List[List[Optional[str]]], file_of_record: List[str], attributes: List[Attribute],
              expected_max_cluster_size: int = 10):
        """RecordsCache.apply (RecordsCache.scala:68-115).  values[r][a] is a string or None (missing)."""
        if len(values) == 0:
            raise ValueError("no records")
        if len(values[0]) != len(attributes):
            raise ValueError("attribute specifications do not match the records")  # :72
        counts = [collections.Counter() for _ in attributes]
        fsz = collections.Counter()
        missing = collections.Counter()
        for rec, f in zip(values, file_of_record):
            fsz[f] += 1
            for a, v in enumerate(rec):
                if v is not None:
                    counts[a][v] += 1
                else:
                    missing[(f, a)] += 1
        indexes = []
        for a, attr in enumerate(attributes):
            vw = {k: float(v) for k, v in counts[a].items()}
            sf = attr.similarity_fn
            indexes.append(AttributeIndex.build(vw, "constant" if sf.is_constant else "levenshtein", sf.threshold,
                                                sf.max_similarity, expected_max_cluster_size))
        fids = sorted(fsz.keys())
        return cls(attributes, indexes, fids, [fsz[f] for f in fids], dict(missing))

    def transform_records(self, values, file_of_record):
        """_transformRecords (RecordsCache.scala:117-133): -> (x int32[R, A], file int32[R])."""
        R, A = len(values), self.num_attributes
        x = np.full((R, A), -1, np.int32)
        for a in range(A):
            ix = self.indexes[a]
            cache = {}
            for r in range(R):
                v = values[r][a]
                if v is None:
                    continue
                i = cache.get(v)
                if i is None:
                    i = ix.value_idx_of(v)
                    cache[v] = i
                x[r, a] = i
        fmap = {f: i for i, f in enumerate(self.file_ids)}
        file = np.array([fmap[f] for f in file_of_record], np.int32)
        return x, file


def read_csv(path, rec_id_col, attribute_names, file_id_col=None, ent_id_col=None, null_value="NA"):
    """Project.scala:173-180 + State.scala:350-371: CSV with header -> record ids, file ids, string values."""
    import gzip

    rec_ids, files, values, ent_ids = [], [], [], []
    opener = (lambda: gzip.open(path, "rt", newline="")) if str(path).endswith(".gz") else (lambda: open(path, newline=""))
    with opener() as fh:
        rd = csv.DictReader(fh)
        for row in rd:
            rec_ids.append(row[rec_id_col])
            files.append(row[file_id_col] if file_id_col else "0")
            values.append([None if (row[a] == null_value or row[a] == "") else row[a] for a in attribute_names])
            if ent_id_col:
                ent_ids.append(row[ent_id_col])
    return rec_ids, files, values, (ent_ids if ent_id_col else None)
