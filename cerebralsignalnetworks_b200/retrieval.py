"""Exact top-k retrieval of encoder embeddings against an image-feature gallery.

Mirrors the faiss calls of the reference's evaluation (utils/Utilities.py:45-58, called from
LstmDistillFromDinoV2Eval.py:333-380): `index = IndexFlatL2(d); index.add(gallery); D, I = index.search(query, k)`,
then the per-class recall / precision bookkeeping of `evaluate` (utils/Utilities.py:60-160).  The search runs in
libcsn_b200 (csn_topk_search: fused distance tiles + warp-level top-k, the [nq, nb] matrix never reaches HBM).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream, call

METRIC_L2, METRIC_IP = 0, 1


def topk_search(gallery: torch.Tensor, query: torch.Tensor, k: int, metric: int = METRIC_L2):
    """gallery [nb, d], query [nq, d] float32 CUDA tensors -> (dist [nq, k] float32, idx [nq, k] int64), best first."""
    _lib.require_gpu()
    for n, t in (("gallery", gallery), ("query", query)):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.dim() != 2:
            raise _lib.CsnError("topk_search: %s must be a contiguous 2-D float32 CUDA tensor" % n)
    nb, d = gallery.shape
    nq, d2 = query.shape
    if d != d2:
        raise _lib.CsnError("topk_search: dimension mismatch (gallery %d, query %d)" % (d, d2))
    nbytes = C.c_size_t(0)
    call("csn_topk_tc_workspace_bytes", nq, nb, d, int(k), C.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=query.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=query.device)
    idx = torch.empty((nq, k), dtype=torch.int64, device=query.device)
    # large galleries: tcgen05 distance GEMM pre-selects 32 candidates per query, exact fp32 re-scoring decides; other
    # shapes run the fused fp32 scan (the entry point routes by shape, same contract either way)
    call("csn_topk_search_tc", _p(gallery), _p(query), nb, nq, d, int(k), int(metric), _p(dist), _p(idx), _p(ws),
         nbytes.value, _stream())
    return dist, idx


class _IndexFlat:
    """The subset of faiss.IndexFlat* the reference uses: attributes d / ntotal / is_trained, add(), search()."""
    metric = METRIC_L2

    def __init__(self, d: int, device="cuda"):
        self.d = int(d)
        self.ntotal = 0
        self.is_trained = True
        self._device = torch.device(device)
        self._chunks = []
        self._gallery = None

    def _to_dev(self, x):
        t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        t = t.reshape(t.shape[0], -1) if t.numel() else t.reshape(t.shape[0], self.d)
        t = t.to(device=self._device, dtype=torch.float32).contiguous()
        if t.shape[1] != self.d:
            raise _lib.CsnError("index of dimension %d got vectors of dimension %d" % (self.d, t.shape[1]))
        return t

    def add(self, x):
        t = self._to_dev(x)
        self._chunks.append(t)
        self._gallery = None
        self.ntotal += t.shape[0]

    def reset(self):
        self._chunks, self._gallery, self.ntotal = [], None, 0

    def search(self, x, k: int):
        """-> (D, I): numpy arrays when the query was numpy (faiss behaviour), CUDA tensors when it was a tensor."""
        if self._gallery is None:
            self._gallery = (torch.cat(self._chunks) if len(self._chunks) != 1 else self._chunks[0]) if self._chunks \
                else torch.empty((0, self.d), dtype=torch.float32, device=self._device)
        q = self._to_dev(x)
        D, I = topk_search(self._gallery, q, k, self.metric)
        if torch.is_tensor(x):
            return D, I
        return D.cpu().numpy(), I.cpu().numpy()


class IndexFlatL2(_IndexFlat):
    """Squared-L2 exact search (faiss.IndexFlatL2, utils/Utilities.py:45)."""
    metric = METRIC_L2


class IndexFlatIP(_IndexFlat):
    """Inner-product exact search (cosine similarity on L2-normalised rows: BASELINE.json config 5)."""
    metric = METRIC_IP


def class_scores(I, gallery_class_ids, query_class_ids, topk=None):
    """Per-class recall / precision exactly as utils/Utilities.py:60-160 accumulates them:
    TP = queries whose class appears among their top-k; Recall = round(100 TP / queries of the class, 2);
    Precision = round(100 * (retrieved rows of the query's class) / (k * queries of the class), 2);
    returns (mean recall, mean precision, per-class dict) with classes in first-appearance order."""
    I = np.asarray(I)
    g = np.asarray(gallery_class_ids)
    q = np.asarray(query_class_ids)
    k = I.shape[1] if topk is None else int(topk)
    hits = (g[I] == q[:, None])
    per_class = {}
    for cls in dict.fromkeys(q.tolist()):
        sel = q == cls
        n = int(sel.sum())
        tp = int(hits[sel].any(axis=1).sum())
        inst = int(hits[sel].sum())
        per_class[cls] = {"TP": tp, "TotalClass": n, "classIntanceRetrival": inst, "TotalRetrival": k * n,
                          "Recall": round((tp * 100) / n, 2), "Precision": round((inst * 100) / (k * n), 2)}
    recall = float(np.array([v["Recall"] for v in per_class.values()]).mean()) if per_class else float("nan")
    precision = float(np.array([v["Precision"] for v in per_class.values()]).mean()) if per_class else float("nan")
    return recall, precision, per_class
