"""Pin the RICES oracle against faiss itself (TEST INFRASTRUCTURE) -- to be run WHERE FAISS IS INSTALLABLE.

faiss is not in this image and not in its offline wheelhouse (`pip download faiss-cpu` finds no distribution; there is no
network), so this script could not be executed during the build and ``oracle/rices.py`` stays "parity UNPINNED".  Anyone with
faiss (``pip install faiss-cpu``) can close that gap:

    python oracle/pin_rices_with_faiss.py        # writes tests/golden/rices_faiss.json

It runs exactly the calls the reference makes -- ``faiss.normalize_L2`` on both matrices, ``IndexFlatIP(D).add(db)``,
``search(q, k)`` (get_question_knn.py:64-76) -- on five seeded cases that include exact ties, an all-zero row, fewer database
rows than k, and a case shaped like the reference's (k = 2048), checks ``oracle/rices.py`` against the result and records
scores and indices.  ``tests/test_rices.py::test_oracle_and_kernel_against_faiss_fixture`` picks the file up when it exists.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rices as orc      # noqa: E402

CASES = [   # name, M, N, D, k, seed
    ("small", 5, 40, 16, 8, 0),
    ("ties_and_zero_row", 6, 64, 32, 16, 1),
    ("fewer_rows_than_k", 3, 10, 8, 16, 2),
    ("clip_like_768", 16, 5000, 768, 64, 3),
    ("k2048", 4, 20000, 64, 2048, 4),
]


def make(name, M, N, D, seed):
    g = np.random.default_rng(seed)
    base = g.standard_normal((1, D)).astype(np.float32)
    db = (base + 0.5 * g.standard_normal((N, D))).astype(np.float32)
    q = (base + 0.5 * g.standard_normal((M, D))).astype(np.float32)
    if name == "ties_and_zero_row":
        db[7] = db[3]
        db[11] = 2.0 * db[3]          # same direction: equal score after normalisation
        db[20] = 0.0
        q[5] = 0.0
    return q, db


def main():
    import faiss
    out = {"faiss": faiss.__version__, "cases": []}
    for name, M, N, D, k, seed in CASES:
        q, db = make(name, M, N, D, seed)
        qn, dbn = q.copy(), db.copy()
        faiss.normalize_L2(qn)
        faiss.normalize_L2(dbn)
        index = faiss.IndexFlatIP(D)
        index.add(dbn)
        Df, If = index.search(qn, k)
        Do, Io = orc.knn_inner_product(q, db, k)
        valid = If >= 0
        assert np.array_equal(valid, Io >= 0), name
        assert np.abs(Df[valid] - Do[valid]).max() < 2e-6, (name, np.abs(Df[valid] - Do[valid]).max())
        # indices: identical except inside groups of (numerically) tied scores, whose order faiss leaves unspecified
        for m in range(M):
            for j in range(k):
                if If[m, j] != Io[m, j]:
                    assert abs(Do[m, j] - Do[m, list(Io[m]).index(If[m, j])]) < 2e-6, (name, m, j)
        out["cases"].append({"name": name, "M": M, "N": N, "D": D, "k": k, "seed": seed,
                             "scores": [[float("%.7g" % x) for x in row] for row in Df.tolist()], "index": If.tolist()})
        print("faiss == oracle on", name)
    path = os.path.join(ROOT, "tests", "golden", "rices_faiss.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, "-- now change the header of oracle/rices.py to PINNED")


if __name__ == "__main__":
    main()
