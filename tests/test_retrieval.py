"""Retrieval row: oracle contract + host scoring on CPU, the fused top-k kernel on the GPU."""
import numpy as np
import pytest
import torch

from oracle.retrieval import evaluate_scores, flat_search


def _data(nq, nb, d, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((nb, d)).astype(np.float32), rng.standard_normal((nq, d)).astype(np.float32)


def test_oracle_flat_search_matches_torch_cdist_topk():
    g, q = _data(37, 500, 48, 0)
    D, I = flat_search(g, q, 5)
    ref = torch.cdist(torch.from_numpy(q).double(), torch.from_numpy(g).double()) ** 2
    rd, ri = ref.topk(5, dim=1, largest=False)
    assert np.array_equal(I, ri.numpy())
    np.testing.assert_allclose(D, rd.numpy(), rtol=1e-10)
    Dip, Iip = flat_search(g, q, 3, metric="ip")
    rs, rj = (torch.from_numpy(q).double() @ torch.from_numpy(g).double().t()).topk(3, dim=1)
    assert np.array_equal(Iip, rj.numpy())
    np.testing.assert_allclose(Dip, rs.numpy(), rtol=1e-10)


def test_oracle_edge_cases():
    g, q = _data(3, 2, 8, 1)
    D, I = flat_search(g, q, 4)           # fewer gallery rows than k
    assert (I[:, 2:] == -1).all() and np.isinf(D[:, 2:]).all() and (I[:, :2] >= 0).all()
    D, I = flat_search(np.zeros((0, 8), np.float32), q, 2)
    assert (I == -1).all()
    g2 = np.repeat(g[:1], 4, axis=0)       # exact ties: lower index first
    _, I = flat_search(g2, q, 3)
    assert (I == np.array([0, 1, 2])).all()


def test_class_scores_follow_the_reference_loop():
    from cerebralsignalnetworks_b200.retrieval import class_scores
    rng = np.random.default_rng(3)
    g_lab = rng.integers(0, 7, 300)
    q_lab = rng.integers(0, 7, 90)
    I = rng.integers(0, 300, (90, 5))
    r0, p0, c0 = evaluate_scores(I, g_lab.tolist(), q_lab.tolist(), 5)
    r1, p1, c1 = class_scores(I, g_lab, q_lab)
    assert r0 == pytest.approx(r1, abs=1e-12) and p0 == pytest.approx(p1, abs=1e-12)
    for cls, v in c0.items():
        for key in ("TP", "TotalClass", "classIntanceRetrival", "TotalRetrival", "Recall", "Precision"):
            assert c1[cls][key] == v[key], (cls, key)


import os

GOLD_SCORES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retrieval_scores.npz")


def test_scoring_matches_the_reference_evaluate_numbers():
    """Recall / precision produced by the reference's OWN Utilities.evaluate (utils/Utilities.py:28-169; only faiss was a
    stand-in when the fixture was generated) against the oracle loop and the product's host-side class_scores."""
    from cerebralsignalnetworks_b200.retrieval import class_scores
    gold = np.load(GOLD_SCORES, allow_pickle=True)
    for i in range(2):
        k = int(gold[f"k{i}"])
        D, I = flat_search(gold[f"gallery{i}"], gold[f"query{i}"], k)
        assert np.array_equal(I, gold[f"I{i}"])
        r0, p0, _ = evaluate_scores(I, gold[f"gallery_ids{i}"].tolist(), gold[f"query_ids{i}"].tolist(), k)
        r1, p1, _ = class_scores(I, gold[f"gallery_ids{i}"], gold[f"query_ids{i}"])
        for r, p in ((r0, p0), (r1, p1)):
            assert r == pytest.approx(float(gold[f"recall{i}"]), abs=1e-9)
            assert p == pytest.approx(float(gold[f"precision{i}"]), abs=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_retrieval_score_fixture_regenerates_from_the_reference():
    from oracle.make_golden import run_reference_evaluate
    gold = np.load(GOLD_SCORES, allow_pickle=True)
    i = 0
    r, p = run_reference_evaluate(gold[f"gallery{i}"], gold[f"query{i}"], gold[f"gallery_ids{i}"], gold[f"query_ids{i}"],
                                  int(gold[f"k{i}"]), int(gold[f"gallery_ids{i}"].max()) + 1)
    assert r == pytest.approx(float(gold[f"recall{i}"]), abs=1e-9) and p == pytest.approx(float(gold[f"precision{i}"]), abs=1e-9)


@pytest.mark.gpu
def test_index_search_reproduces_the_reference_evaluate_numbers():
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200.retrieval import class_scores
    gold = np.load(GOLD_SCORES, allow_pickle=True)
    for i in range(2):
        index = csn.IndexFlatL2(gold[f"gallery{i}"].shape[1])
        index.add(gold[f"gallery{i}"])
        _, I = index.search(gold[f"query{i}"], int(gold[f"k{i}"]))
        r, p, _ = class_scores(I, gold[f"gallery_ids{i}"], gold[f"query_ids{i}"])
        assert r == pytest.approx(float(gold[f"recall{i}"]), abs=1e-9) and p == pytest.approx(float(gold[f"precision{i}"]), abs=1e-9)


def _check_against_oracle(g, q, k, metric):
    import cerebralsignalnetworks_b200 as csn
    index = (csn.IndexFlatL2 if metric == "l2" else csn.IndexFlatIP)(g.shape[1])
    if len(g):
        index.add(g[: len(g) // 2])
        index.add(g[len(g) // 2:])      # incremental add, as faiss allows
    assert index.ntotal == len(g) and index.is_trained
    D, I = index.search(q, k)
    assert isinstance(D, np.ndarray) and D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (len(q), k)
    Do, Io = flat_search(g, q, k, metric)
    valid = Io >= 0
    # a row is decidable in fp32 when the oracle's neighbouring scores (ranks 1..k+1) are separated
    Dk, _ = flat_search(g, q, min(k + 1, max(len(g), 1)), metric) if len(g) else (Do, Io)
    gaps = np.abs(np.diff(Dk, axis=1))
    scale = np.maximum(np.abs(Dk[:, :1]), 1.0)
    decidable = (np.nan_to_num(gaps, nan=np.inf, posinf=np.inf) > 1e-5 * scale).all(axis=1) if gaps.size else np.ones(len(q), bool)
    assert decidable.any() or len(g) < 2 * k
    assert np.array_equal(I[decidable], Io[decidable])
    # rows with near-ties: the same SET up to swaps among neighbours whose float64 scores differ by < 1e-5 relative
    for r in np.nonzero(~decidable)[0]:
        swapped = I[r] != Io[r]
        if swapped.any():
            true = ((g[I[r][swapped]].astype(np.float64) - q[r]) ** 2).sum(1) if metric == "l2" else g[I[r][swapped]].astype(np.float64) @ q[r]
            np.testing.assert_allclose(true, Do[r][swapped], rtol=3e-5, atol=3e-5)
    assert (I[~valid] == -1).all()
    np.testing.assert_allclose(D[valid], Do[valid], rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("nq,nb,d,k", [(1, 1, 1, 1), (5, 3, 7, 5), (33, 129, 32, 5), (64, 1000, 100, 8), (200, 10000, 384, 5),
                                       (70, 4097, 768, 32), (3, 0, 16, 2)])
def test_topk_search_matches_oracle(metric, nq, nb, d, k):
    g, q = _data(nq, nb, d, 10 + nq)
    if metric == "ip":
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    _check_against_oracle(g, q, k, metric)


@pytest.mark.gpu
def test_topk_ties_and_self_retrieval():
    import cerebralsignalnetworks_b200 as csn
    g, _ = _data(1, 600, 24, 5)
    g[300:] = g[:300]                       # every row has an exact duplicate 300 places later
    index = csn.IndexFlatL2(24)
    index.add(g)
    D, I = index.search(g[:50], 2)          # "sanity check" of Utilities.py:52: a row finds itself first
    assert np.array_equal(I[:, 0], np.arange(50)) and np.array_equal(I[:, 1], np.arange(50) + 300)
    assert (D == 0).all()
    Dt, It = index.search(torch.from_numpy(g[:4]).cuda(), 2)
    assert torch.is_tensor(Dt) and It.is_cuda and np.array_equal(It.cpu().numpy(), I[:4])


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("nq,nb,d,k", [(40, 40000, 64, 10), (9, 2048, 8, 16), (300, 5000, 136, 1)])
def test_tensor_core_preselection_matches_oracle(metric, nq, nb, d, k):
    """Galleries the tcgen05 pre-selection path serves (csn_topk_search_tc): two gallery chunks, the smallest shapes it
    takes, a query count that is not a multiple of the GEMM tile."""
    g, q = _data(nq, nb, d, 77 + nq)
    if metric == "ip":
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    _check_against_oracle(g, q, k, metric)


@pytest.mark.gpu
def test_tensor_core_path_duplicates_and_agreement_with_the_scan_kernel():
    """Exact duplicates in a large gallery: the exact re-scoring reports distance 0 and orders ties by index; and the
    two product kernels (GEMM pre-selection + re-scoring, fused fp32 scan) return the same neighbours."""
    import ctypes as C
    import cerebralsignalnetworks_b200 as csn
    from cerebralsignalnetworks_b200 import _lib
    g, _ = _data(1, 6000, 32, 9)
    g[3000:] = g[:3000]
    index = csn.IndexFlatL2(32)
    index.add(g)
    D, I = index.search(g[:64], 2)
    assert np.array_equal(I[:, 0], np.arange(64)) and np.array_equal(I[:, 1], np.arange(64) + 3000)
    assert (D == 0).all()
    gal, qry = torch.from_numpy(g).cuda(), torch.from_numpy(_data(50, 1, 32, 4)[1]).cuda()
    Dt, It = csn.retrieval.topk_search(gal, qry, 5)
    nbytes = C.c_size_t(0)
    _lib.call("csn_topk_workspace_bytes", 50, 6000, 5, C.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    Ds, Is = torch.empty(50, 5, device="cuda"), torch.empty(50, 5, dtype=torch.int64, device="cuda")
    _lib.call("csn_topk_search", gal.data_ptr(), qry.data_ptr(), 6000, 50, 32, 5, 0, Ds.data_ptr(), Is.data_ptr(), ws.data_ptr(), None)
    torch.cuda.synchronize()
    assert torch.equal(It, Is)
    np.testing.assert_allclose(Dt.cpu().numpy(), Ds.cpu().numpy(), rtol=2e-6, atol=1e-6)


@pytest.mark.gpu
def test_topk_rejects_bad_arguments():
    import cerebralsignalnetworks_b200 as csn
    index = csn.IndexFlatL2(8)
    index.add(np.zeros((4, 8), np.float32))
    with pytest.raises(csn.CsnError):
        index.search(np.zeros((2, 8), np.float32), 33)
    with pytest.raises(csn.CsnError):
        index.search(np.zeros((2, 9), np.float32), 1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 4e-2)])
def test_encode_trials_63_channels_then_retrieve(dtype, tol):
    """BASELINE config 5 in miniature: 63-channel trials (not a multiple of 8 -> zero-channel padding on the bf16
    path) -> band-pass -> LSTM inference -> embeddings -> exact top-k against a gallery; embeddings are compared with
    the torch-CPU restatement of the encoder on scipy-filtered input, retrieved indices with the oracle search."""
    import cerebralsignalnetworks_b200 as csn
    from oracle.filters import design_bandpass_sos, sosfilt_np
    B, Cc, T, H, Dd = 24, 63, 256, 64, 48
    torch.manual_seed(7)
    model = csn.Model(Cc, H, 2, Dd, include_top=False, compute_dtype=dtype).cuda().eval()
    sos = design_bandpass_sos(5.0, 95.0, 1000.0, 4)
    rng = np.random.default_rng(2)
    eeg = rng.standard_normal((B, Cc, T)).astype(np.float32)
    emb = model.encode_trials(torch.from_numpy(eeg).cuda(), sos=sos).cpu()
    assert emb.shape == (B, Dd)
    # reference: scipy sosfilt (float64) -> torch.nn.LSTM on the CPU -> Linear
    filt = torch.from_numpy(sosfilt_np(sos, eeg).astype(np.float32)).permute(0, 2, 1).contiguous()  # [B, T, C]
    ref_lstm = torch.nn.LSTM(Cc, H, 2, batch_first=True)
    with torch.no_grad():
        for n, p in model.lstm.named_parameters():
            getattr(ref_lstm, n).copy_(p.cpu())
        out, _ = ref_lstm(filt)
        ref = out[:, -1] @ model.output.weight.cpu().t() + model.output.bias.cpu()
    assert (emb - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    gallery = rng.standard_normal((500, Dd)).astype(np.float32)
    index = csn.IndexFlatL2(Dd)
    index.add(gallery)
    D, I = index.search(emb.numpy(), 5)
    Do, Io = flat_search(gallery, emb.numpy(), 5)
    assert np.array_equal(I, Io)
    np.testing.assert_allclose(D, Do, rtol=2e-5, atol=2e-5)
