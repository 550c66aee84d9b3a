"""Library-compatibility patches for running the 2020-era reference on the 2026 software stack of this image
(pandas 3 removed DataFrame.append, sklearn renamed ``affinity``, networkx 3 removed to_numpy_matrix).
The reference sources are NOT touched; these restore the removed library entry points."""
import os

if os.environ.get("FDB_REF_SHIMS") == "1":
    try:
        import pandas as pd

        if not hasattr(pd.DataFrame, "append"):
            def _append(self, other, ignore_index=False, **kw):
                if len(self.columns) == 0 or len(self) == 0:
                    out = other.copy()
                    return out.reset_index(drop=True) if ignore_index else out
                return pd.concat([self, other], ignore_index=ignore_index)

            pd.DataFrame.append = _append
    except Exception:
        pass
    try:
        import sklearn.cluster as _sc

        class _AgglomerativeClustering(_sc.AgglomerativeClustering):
            def __init__(self, n_clusters=2, *, affinity=None, metric="euclidean", memory=None, connectivity=None,
                         compute_full_tree="auto", linkage="ward", distance_threshold=None, compute_distances=False):
                super().__init__(n_clusters=n_clusters, metric=affinity if affinity is not None else metric,
                                 memory=memory, connectivity=connectivity, compute_full_tree=compute_full_tree,
                                 linkage=linkage, distance_threshold=distance_threshold,
                                 compute_distances=compute_distances)
                self.affinity = affinity

        _sc.AgglomerativeClustering = _AgglomerativeClustering
    except Exception:
        pass
    try:
        import networkx as nx

        if not hasattr(nx, "to_numpy_matrix"):
            nx.to_numpy_matrix = nx.to_numpy_array
    except Exception:
        pass
