"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the retrieval step of the reference's evaluation, utils/Utilities.py:28-169 (`evaluate`):
faiss.IndexFlatL2(d).add(gallery); D, I = index.search(query, k)  (lines 45-58), then the per-class TP / recall /
precision bookkeeping (lines 60-160).  faiss is an un-vendored third-party dependency that is absent from this image
(no version pin in the reference): its published IndexFlatL2 contract is restated -- exhaustive squared-L2 search,
results sorted by increasing distance -- in float64, so the ranking is the exact one.  PARITY UNPINNED by the reference
(it holds no fixtures for this step); the scoring loop below follows the reference line by line.
"""
from __future__ import annotations

import numpy as np


def flat_search(gallery, query, k, metric="l2"):
    """-> (D [nq, k] float64, I [nq, k] int64).  l2: squared distances ascending; ip: inner products descending.
    Ties go to the lower gallery index (stable sort); missing neighbours (nb < k) are index -1."""
    q = np.asarray(query, dtype=np.float64)
    q = q.reshape(len(q), -1)
    g = np.asarray(gallery, dtype=np.float64).reshape(len(gallery), q.shape[1])
    nq, nb = q.shape[0], g.shape[0]
    D = np.full((nq, k), np.inf if metric == "l2" else -np.inf)
    I = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        if nb == 0:
            continue
        if metric == "l2":
            s = ((g - q[i]) ** 2).sum(axis=1)
            order = np.argsort(s, kind="stable")[:k]
        else:
            s = g @ q[i]
            order = np.argsort(-s, kind="stable")[:k]
        D[i, :len(order)] = s[order]
        I[i, :len(order)] = order
    return D, I


def evaluate_scores(I, gallery_labels, query_labels, topK):
    """The accumulation loop of utils/Utilities.py:60-160 on plain class ids (ClassName == ClassId here).
    Returns (Recall_Total, Precision_Total, class_scores)."""
    class_scores = {}
    for query_idx, search_res in enumerate(I):
        test_label = query_labels[query_idx]
        if test_label not in class_scores:  # :83-99
            class_scores[test_label] = {"TP": 0, "classIntanceRetrival": 0, "TotalRetrival": 0, "TotalClass": 0,
                                        "Recall": "", "Precision": ""}
        retrieved = [gallery_labels[j] for j in search_res]  # :101-106
        count = sum(1 for r in retrieved if r == test_label)  # :109-116 (np.unique counts of the query's class)
        if test_label in retrieved:  # :122-125
            class_scores[test_label]["TP"] += 1
            class_scores[test_label]["classIntanceRetrival"] += count
        class_scores[test_label]["TotalRetrival"] += topK  # :130
        class_scores[test_label]["TotalClass"] += 1        # :131
        c = class_scores[test_label]
        c["Recall"] = round(((c["TP"] * 100) / c["TotalClass"]), 2)                        # :145
        c["Precision"] = round(((c["classIntanceRetrival"] * 100) / c["TotalRetrival"]), 2)  # :146
    recall = np.array([c["Recall"] for c in class_scores.values()]).mean()       # :149-157
    precision = np.array([c["Precision"] for c in class_scores.values()]).mean()
    return recall, precision, class_scores
