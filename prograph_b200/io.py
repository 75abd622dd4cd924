"""Persistence of a graph: the reference's pickle (utils/save.py:5-39) plus a CSR sidecar.

The pickle keeps the reference's format -- the frame without the ``Tokenized`` column, one
``(indices, weights)`` tuple per row in ``Neighbours`` -- so files stay interchangeable.  Next to it
``<name>.graph.npz`` stores the same adjacency as three flat arrays (indptr, indices, weights): a
reload attaches them to the frame's column so that degree / adjacency / laplacian need no per-row
loop over a million tiny arrays.
"""
import hashlib
import os

import numpy as np


def graph_fingerprint(column):
    """Digest of a ``Neighbours`` column (one (indices, weights) tuple per row): degree, index and
    weight bytes of every row.  The sidecar carries it, so a pickle rewritten by anything else (the
    reference's own ``save``, an edited frame) never picks up a stale sidecar."""
    h = hashlib.blake2b(digest_size=16)
    for idx, w in column:
        idx, w = np.asarray(idx), np.asarray(w)
        h.update(np.int64(len(idx)).tobytes())
        h.update(np.ascontiguousarray(idx, dtype=np.int64).tobytes())
        h.update(str(w.dtype).encode())
        h.update(np.ascontiguousarray(w).tobytes())
    return h.hexdigest()


def sidecar_path(pickle_path):
    return os.path.splitext(pickle_path)[0] + ".graph.npz"


def save(pgraph, name=None, ext=".pkl", directory=None, ignored_cols=["Tokenized"]):
    """Same arguments and behaviour as the reference's ``save`` (utils/save.py:5-39): pickles the
    frame minus ``ignored_cols`` to ``directory + name + ext``; additionally writes the CSR sidecar."""
    if directory is None:
        if hasattr(pgraph, "file") and "/" in str(pgraph.file):
            directory, file = pgraph.file.rsplit("/", 1)
            directory += "/"
        else:
            directory, file = "./", getattr(pgraph, "file", "pgraph")
    else:
        file = os.path.basename(str(getattr(pgraph, "file", "pgraph")))
    if not name:
        name = file.rsplit(".", 1)[0] + "_pgraph"
    path = directory + name + ext
    print(f"Saving Graph to {name + ext}")
    try:
        pgraph.graph[[c for c in pgraph.graph if c not in ignored_cols]].to_pickle(path)
        if "Neighbours" in pgraph.graph:
            t = pgraph._table_for("Neighbours")
            np.savez(sidecar_path(path), indptr=t.indptr, idx=t.idx, w=t.w,
                     fingerprint=np.array(graph_fingerprint(pgraph.graph["Neighbours"])))
    except Exception as e:          # the reference reports and carries on
        print("Error occurred during saving:", e)
    return True


def attach_sidecar(pgraph, pickle_path):
    """After loading a pickle: if its sidecar is there and matches the frame, register the flat
    CSR arrays for the ``Neighbours`` column."""
    from .graph import NeighbourTable
    path = sidecar_path(pickle_path)
    if not os.path.exists(path) or "Neighbours" not in pgraph.graph or len(pgraph.graph) == 0:
        return False
    with np.load(path) as z:
        if "fingerprint" not in z:
            return False
        table = NeighbourTable(z["indptr"], z["idx"], z["w"])
        fingerprint = str(z["fingerprint"])
    col = pgraph.graph["Neighbours"]
    if table.n_rows != len(col) or fingerprint != graph_fingerprint(col):
        return False                 # stale sidecar: the exports flatten the column instead
    pgraph._remember(table, [col.iloc[0]])
    return True
