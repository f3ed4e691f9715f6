"""vector-store_b200 — B200-native (sm_100a) ANN engine for the USearch index path of
scylladb/vector-store.  The product is `libvsb200.so` (CUDA + C ABI, include/vsb200.h); this
package is the thin host-side mirror of the reference's index layer used by tests and bench.

No CPU fallback exists: every search/add goes through the CUDA library or raises.
"""
from .host.native import VsbError, lib, lib_path, version  # noqa: F401
from .host.index import Batcher, GpuIndex, IndexSet, Metric, Scalar  # noqa: F401
from .host.distance import Distance, SimilarityScore, SpaceType  # noqa: F401
from .host.actor import (IndexActor, Quantization, VsIndexConfiguration, WrongEmbeddingDimension,  # noqa: F401
                         new_index_factory_gpu)
