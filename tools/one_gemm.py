"""Run one GEMM shape of the step a few times (target for `ncu --set full`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eavqa_b200 import lib
L = lib.load()
M, N, K = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "gelu2"
code = int(sys.argv[5]) if len(sys.argv) > 5 else 0
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
if mode == "gelu2":
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); out2 = torch.empty_like(out)
    go = lambda: lib.check(L.eavqa_op_gemm(A.data_ptr(), K, B.data_ptr(), K, M, N, K, out.data_ptr(), N, 0, bias.data_ptr(), None, 0, 1, None, 0, 0, out2.data_ptr(), N, code, st))
elif mode == "res":
    out = torch.empty(M, N, device="cuda"); R = torch.randn(M, N, device="cuda")
    go = lambda: lib.check(L.eavqa_op_gemm(A.data_ptr(), K, B.data_ptr(), K, M, N, K, out.data_ptr(), N, 1, bias.data_ptr(), R.data_ptr(), N, 0, None, 0, 0, None, 0, code, st))
else:
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    go = lambda: lib.check(L.eavqa_op_gemm(A.data_ptr(), K, B.data_ptr(), K, M, N, K, out.data_ptr(), N, 0, None, None, 0, 0, None, 0, 0, None, 0, code, st))
for _ in range(5):
    go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    go()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{M}x{N}x{K} {mode} code={code}: {ms*1e3:.1f} us, {2.0*M*N*K/ms/1e9:.0f} TFLOP/s")
