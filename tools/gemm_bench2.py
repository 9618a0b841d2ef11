"""Sub-wave GEMM study: tile width sweep with the residual epilogue, cold vs back-to-back timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200 import lib

L = lib.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def make(M, N, K, mode):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    fp32 = mode in ("res", "f32")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    R = torch.randn(M, N, device="cuda") if mode == "res" else None
    bias = torch.randn(N, device="cuda")
    def go(bn):
        lib.check(L.eavqa_op_gemm(A.data_ptr(), K, B.data_ptr(), K, M, N, K, out.data_ptr(), N, int(fp32), bias.data_ptr(),
                                  R.data_ptr() if R is not None else None, N if R is not None else 0, 0, None, 0, 0, None, 0,
                                  bn + 1000, st))
    return go


def timeit(go, bn, cold=True, reps=9, burst=1):
    for _ in range(3):
        go(bn)
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(burst):
            go(bn)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / burst)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


shapes = [(128, 128, 64, "bf16"), (128, 256, 768, "bf16"), (5120, 768, 768, "bf16"), (5120, 768, 768, "res"), (5120, 768, 1536, "res"),
          (5120, 1536, 768, "bf16"), (5120, 2304, 768, "bf16"), (5120, 768, 2304, "bf16"), (768, 768, 5120, "f32"),
          (1536, 768, 5120, "f32"), (2304, 768, 5120, "f32"), (12800, 768, 768, "bf16"), (12800, 768, 768, "res"),
          (12800, 768, 3072, "res"), (256, 7680, 512, "f32")]
print("microseconds per launch: cold (L2 flushed, single) / warm back-to-back x20")
for (M, N, K, mode) in shapes:
    go = make(M, N, K, mode)
    row = []
    for bn in (64, 128, 192, 256):
        row.append(f"bn{bn}: {timeit(go, bn):6.1f}/{timeit(go, bn, cold=False, burst=20):6.1f}")
    ideal = 2.0 * M * N * K / 1.4e15 * 1e6
    print(f"M={M:5d} N={N:5d} K={K:5d} {mode:4s} ideal@1400TF {ideal:5.1f}us | " + "  ".join(row), flush=True)
