"""Standalone timing of the tcgen05 GEMM over tile / cluster shapes (CUDA events, L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200 import lib

L = lib.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def run(M, N, K, bn, cl, fp32=False, res=False, reps=10):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    R = torch.randn(M, N, device="cuda") if res else None
    st = torch.cuda.current_stream().cuda_stream
    def go():
        lib.check(L.eavqa_op_gemm(A.data_ptr(), K, B.data_ptr(), K, M, N, K, out.data_ptr(), N, int(fp32), None,
                                  R.data_ptr() if res else None, N if res else 0, 0, None, 0, 0, None, 0, bn + 1000 * cl, st))
    for _ in range(3):
        go()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    return ms, 2.0 * M * N * K / ms / 1e9


shapes = [(12800, 768, 768), (12800, 2304, 768), (12800, 3072, 768), (12800, 768, 3072), (10240, 50304, 768), (10240, 768, 50304),
          (5120, 768, 768), (5120, 2304, 768), (2304, 768, 5120), (8192, 8192, 8192)]
for (M, N, K) in shapes:
    row = []
    for bn in (128, 192, 256):
        for cl in (1, 8):
            try:
                ms, tf = run(M, N, K, bn, cl)
                row.append(f"bn{bn}/c{cl}:{tf:6.0f}")
            except Exception as e:
                row.append(f"bn{bn}/c{cl}: ERR")
    print(f"M={M} N={N} K={K}  TFLOP/s  " + "  ".join(row), flush=True)
