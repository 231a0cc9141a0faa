"""Probe of the GPU box's host side: fp32 -> bf16 conversion rate on the host cores vs PCIe H2D rate."""
import os, time, torch
n = os.cpu_count()
print("cpus", n, "torch threads", torch.get_num_threads())
B, T, F = 1024, 80, 4096
src = torch.randn(B, T, F).pin_memory()
dst = torch.empty(B, T, F, dtype=torch.bfloat16).pin_memory()
for th in (4, 8, n):
    torch.set_num_threads(th)
    dst.copy_(src)
    t0 = time.time()
    for _ in range(3):
        dst.copy_(src)
    dt = (time.time() - t0) / 3
    print(f"convert fp32->bf16 threads={th}: {dt*1e3:.1f} ms  read {src.numel()*4/dt/1e9:.1f} GB/s")
dev = torch.device("cuda")
d32 = torch.empty(B, T, F, device=dev)
d16 = torch.empty(B, T, F, dtype=torch.bfloat16, device=dev)
for name, h, d in (("fp32", src, d32), ("bf16", dst, d16)):
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    print(f"H2D {name}: {dt*1e3:.1f} ms  {h.numel()*h.element_size()/dt/1e9:.1f} GB/s")
# overlapped: convert chunk i+1 on the host while chunk i is in flight
torch.set_num_threads(n)
chunks = 8
cs = B // chunks
t0 = time.time()
for _ in range(3):
    for i in range(chunks):
        dst[i*cs:(i+1)*cs].copy_(src[i*cs:(i+1)*cs])
        d16[i*cs:(i+1)*cs].copy_(dst[i*cs:(i+1)*cs], non_blocking=True)
    torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print(f"pipelined convert+H2D bf16: {dt*1e3:.1f} ms per 1024 videos -> {B/dt:.0f} videos/s")
