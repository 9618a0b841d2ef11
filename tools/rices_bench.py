"""RICES retrieval at the reference's scale: VQA2 train questions as the database (443 757 x 768 CLIP ViT-L/14 text
embeddings), k = 2048 (get_question_knn.py:73), a block of val questions as queries.  CUDA events; the database is
re-normalised / re-packed inside every call (as a faiss index build + search would be).

    python tools/rices_bench.py [n_queries] [n_database] [reps]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200.rices import knn_inner_product


def run(M=4096, N=443757, D=768, k=2048, reps=10):
    g = torch.Generator(device="cuda").manual_seed(0)
    base = torch.randn(1, D, device="cuda", generator=g)
    db = base + 0.5 * torch.randn(N, D, device="cuda", generator=g)
    q = base + 0.5 * torch.randn(M, D, device="cuda", generator=g)
    for _ in range(2):                                     # warm-up WITH THE TIMED SHAPE: kernel attributes, and the cached
        knn_inner_product(q, db, k)                        # workspace reaches its final size (no allocation while timing)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        scores, index = knn_inner_product(q, db, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * M * N * D
    return {"metric": "rices_knn_queries_per_sec", "value": M / (ms * 1e-3), "unit": "queries/s", "ms_per_call": ms,
            "config": {"workload": "faiss.normalize_L2 + IndexFlatIP.search of get_question_knn.py:64-76 on synthetic CLIP-like "
                                   "embeddings", "n_queries": M, "n_database": N, "dim": D, "k": k},
            "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "executed_tflops": 3 * flop / (ms * 1e-3) / 1e12,
            "top1_score": float(scores[:, 0].mean())}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    print(run(*(a[:2] if len(a) >= 2 else a), **({"reps": a[2]} if len(a) > 2 else {})))
