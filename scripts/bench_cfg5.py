"""cfg 5 of BASELINE.json on one GPU: Perils-shaped trials (63 channels x 2000 samples) -> band-pass -> LSTM encoder
inference -> embeddings -> exact top-k over a 10 k-image gallery (L2 as the reference's faiss index, and cosine)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cerebralsignalnetworks_b200 as csn
from cerebralsignalnetworks_b200 import retrieval
B, C, T, H, D, NB, K = int(os.environ.get("PB", "256")), 63, 2000, 128, 384, 10000, 5
torch.manual_seed(43)
model = csn.Model(C, H, 1, D, include_top=False, compute_dtype=torch.bfloat16).cuda().eval()
sos = csn.EEGFilters(1000.0).sos(5.0, 95.0, 4)
g = torch.Generator(device="cuda").manual_seed(1)
eeg = [torch.randn(B, C, T, device="cuda", generator=g) for _ in range(3)]
gallery = torch.randn(NB, D, device="cuda", generator=g)
def timed(fn, reps):
    fn(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        out = fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
ms_enc, emb = timed(lambda i: model.encode_trials(eeg[i % 3], sos=sos), 6)
print(f"encoder inference B={B} 63ch x 2000: {ms_enc:.2f} ms  {B / ms_enc * 1e3:.0f} trials/s")
for name, metric in (("L2", retrieval.METRIC_L2), ("cosine", retrieval.METRIC_IP)):
    q, gal = emb, gallery
    if metric == retrieval.METRIC_IP:
        q = torch.nn.functional.normalize(emb); gal = torch.nn.functional.normalize(gallery)
    ms, (Dd, I) = timed(lambda i: retrieval.topk_search(gal, q, K, metric), 20)
    flop = 2.0 * B * NB * D * (1.5 if metric == retrieval.METRIC_L2 else 1.0)
    print(f"top-{K} {name} over {NB} x {D}: {ms * 1e3:.1f} us  {B / ms * 1e3:.0f} queries/s  {flop / ms / 1e9:.2f} TFLOP/s fp32")
big = torch.randn(2048, D, device="cuda", generator=g)
ms, _ = timed(lambda i: retrieval.topk_search(gallery, big, K, retrieval.METRIC_L2), 10)
print(f"top-{K} L2, 2048 queries: {ms * 1e3:.1f} us  {2048 / ms * 1e3:.0f} queries/s  {3.0 * 2048 * NB * D / ms / 1e9:.2f} TFLOP/s fp32")
