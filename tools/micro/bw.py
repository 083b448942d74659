"""HBM bandwidth by access mix (torch ops as the yardstick): copy 1:1, write-only, read-only, int16->f32 (1:2)."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / 1e3
N = 1 << 30
a = torch.empty(N, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
s = t(lambda: b.copy_(a)); print(f"copy bf16 1Gi (r+w 4 GiB)      : {4*N/s/1e9:8.1f} GB/s")
s = t(lambda: b.zero_()); print(f"memset 2 GiB (write only)      : {2*N/s/1e9:8.1f} GB/s")
s = t(lambda: a.sum()); print(f"sum 2 GiB (read only)          : {2*N/s/1e9:8.1f} GB/s")
i16 = torch.empty(N // 2, dtype=torch.int16, device="cuda").random_(-3000, 3000)
f32 = torch.empty(N // 2, dtype=torch.float32, device="cuda")
s = t(lambda: f32.copy_(i16)); print(f"int16->f32 512Mi (r 1 GiB + w 2 GiB): {3*(N//2)*2/s/1e9:8.1f} GB/s")
f = torch.empty(N // 2, dtype=torch.float32, device="cuda")
s = t(lambda: f.copy_(f32)); print(f"copy f32 512Mi (r+w 4 GiB)     : {8*(N//2)/s/1e9:8.1f} GB/s")
