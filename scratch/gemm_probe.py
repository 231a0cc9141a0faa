"""Time the persistent store-epilogue GEMM (vc_linear, bf16) at the decoder-LSTM shapes: how fast is the main loop alone?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_captioning_b200 import _native
torch.manual_seed(0)
for (M, N, K) in ((5120, 2048, 1536), (5120, 2048, 1024), (5120, 10000, 512), (81920, 4096, 512), (10240, 2048, 1536), (9472, 2048, 1536)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    # vc_linear converts A/W to bf16 first (cast kernels) -- time the whole call and subtract a cast-only estimate
    for _ in range(3): C = _native.linear(A, W, b, precision="bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): C = _native.linear(A, W, b, precision="bf16")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    cast_bytes = (M * K + N * K) * 6
    cast_ms = cast_bytes / 6.0e9 * 1e-0 / 1e3 * 1e3 / 1e3
    print(f"M={M} N={N} K={K}: {ms*1e3:.1f} us per call incl. casts (~{cast_bytes/6.2e12*1e6:.1f} us of cast traffic), {2*M*N*K/ms/1e9:.0f} TF/s raw")
