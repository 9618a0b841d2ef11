"""Single-token decode GEMMs (M = batch = 128 rows): time per launch of the tcgen05 GEMM with and without split-K,
weights rotated through 24 buffers so that they stream from HBM as in the real decode loop (CUDA events, 48 launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200 import lib
L = lib.load()
st = torch.cuda.current_stream().cuda_stream
M = 128
for (N, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (50304, 1024)):
    nbuf = 24 if N < 10000 else 3
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Bs = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    out = torch.zeros(M, N, device="cuda")
    for bn, split in ((64, 1), (64, 2), (64, 4), (64, 8), (128, 1), (128, 4), (128, 8), (128, 16)):
        if split > K // 64:
            continue
        code = bn + 1000 * 1 + 100000 * split
        go = lambda i: lib.check(L.eavqa_op_gemm(A.data_ptr(), K, Bs[i % nbuf].data_ptr(), K, M, N, K, out.data_ptr(), N, 1, None, None, 0, 0,
                                                 None, 0, 0, None, 0, code, st))
        for i in range(nbuf):
            go(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 48
        for i in range(n):
            go(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        tiles = (N + bn - 1) // bn * split
        print(f"M=128 N={N} K={K} bn={bn} split={split} ctas={min(tiles,148)}: {us:.1f} us/launch, weights {N*K*2/us/1e3:.0f} GB/s")
