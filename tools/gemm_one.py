"""One GEMM shape, a few launches (for ncu): python tools/gemm_one.py M N K [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagined_speech_translation_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
it = int(sys.argv[4]) if len(sys.argv) > 4 else 5
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(it):
    ops.gemm(a, b, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(it):
    ops.gemm(a, b, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / it
print(f"{M}x{N}x{K}: {t*1e3:.1f} us  {2.0*M*N*K/t/1e9:.1f} TF/s")
