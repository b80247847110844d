"""Times the tcgen05 GEMM on encoder-shaped problems (CUDA events), next to torch.matmul."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagined_speech_translation_b200 import ops

def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

shapes = [(9472, 768, 768), (9472, 3072, 768), (9472, 768, 3072), (9472, 1536, 768), (9472, 2304, 768),
          (8448, 128, 18576), (8192, 8192, 8192), (4096, 51271 // 8 * 8, 768), (37888, 768, 768), (37888, 3072, 768)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t_ours = timeit(lambda: ops.gemm(a, b, out=out))
    t_ref = timeit(lambda: torch.matmul(a, b.t(), out=out))
    fl = 2.0 * M * N * K
    line = f"M={M:6d} N={N:6d} K={K:6d}  ours {t_ours*1e3:8.1f} us {fl/t_ours/1e9:7.1f} TF/s | cublas {t_ref*1e3:8.1f} us {fl/t_ref/1e9:7.1f} TF/s"
    for bn in (128, 256):
        t = timeit(lambda: ops.gemm(a, b, out=out, force_block_n=bn))
        line += f" | bn{bn} {fl/t/1e9:6.1f}"
    print(line, flush=True)
# wgrad / dgrad forms
M, N, K = 768, 3072, 9472
dy = torch.randn(K, M, device="cuda").bfloat16(); x = torch.randn(K, N, device="cuda").bfloat16()
o = torch.empty(M, N, device="cuda")
t = timeit(lambda: ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=o))
print(f"wgrad {M}x{N}x{K}: {t*1e3:.1f} us {2.0*M*N*K/t/1e9:.1f} TF/s")
M, N, K = 9472, 768, 3072
dy = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(K, N, device="cuda").bfloat16()
o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
t = timeit(lambda: ops.gemm(dy, w, b_mn_major=True, out=o))
print(f"dgrad {M}x{N}x{K}: {t*1e3:.1f} us {2.0*M*N*K/t/1e9:.1f} TF/s")
