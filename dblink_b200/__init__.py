"""dblink_b200: B200-native Gibbs sweep for cleanzr/dblink's record-linkage model.

The numerical work lives in libdblink_b200.so (CUDA, sm_100a) behind the C ABI of include/dblink_b200.h;
this package is the thin host layer that mirrors the reference's objects around the sweep.
"""
from ._lib import SAMPLERS, PCG_I, PCG_II, GIBBS, GIBBS_SEQ  # noqa: F401
from .engine import AttributeIndex, KDTreePartitioner, GibbsEngine, DblinkError, similarity  # noqa: F401
from .records import Attribute, SimilarityFn, RecordsCache, read_csv  # noqa: F401

__version__ = "0.1.0"
