"""Diagnostics for the tcgen05 GEMM bring-up: runs tiny identity / one-hot problems and prints how the output is
permuted when it is wrong.  Writes nothing; run under `timeout` on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200 import lib

L = lib.load()


def gemm(A, B, bn):
    M, K = A.shape
    N = B.shape[0]
    out = torch.full((M, N), -777.0, device="cuda")
    lib.check(L.eavqa_op_gemm(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, out.data_ptr(), N, 1, None, None, 0, 0,
                              None, 0, 0, None, 0, bn, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 64, 128), (128, 128, 128, 128), (256, 256, 256, 256), (128, 192, 192, 192)]:
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).to("cuda", torch.bfloat16)
    B = torch.zeros(N, K, device="cuda", dtype=torch.bfloat16)
    idx = torch.arange(min(N, K), device="cuda")
    B[idx, idx] = 1
    out = gemm(A, B, bn)
    ref = A.float() @ B.float().t()
    ok = torch.equal(out, ref)
    print(f"M={M} N={N} K={K} bn={bn}: identity {'OK' if ok else 'MISMATCH'}; untouched={(out == -777).float().mean().item():.3f} "
          f"maxerr={(out - ref).abs().max().item():.4f}")
    if not ok:
        # for a few output elements find where their value came from in A
        Af = A.float()
        for r in (0, 1, 8, 33, 127):
            for c in (0, 1, 8, 17, 63):
                if c >= min(N, K):
                    continue
                v = out[r, c].item()
                hits = (Af == v).nonzero()[:3].tolist()
                print(f"   out[{r},{c}]={v:.4f} expected A[{r},{c}]={Af[r, c].item():.4f}; value found in A at {hits}")
    g2 = torch.Generator().manual_seed(1)
    B2 = torch.randn(N, K, generator=g2).to("cuda", torch.bfloat16)
    out = gemm(A, B2, bn)
    ref = A.float() @ B2.float().t()
    print(f"   random: maxerr={(out - ref).abs().max().item():.5f} refmax={ref.abs().max().item():.3f}")
print("launches", L.eavqa_launch_count())
