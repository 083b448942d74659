import torch, time
n = 635_040_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for pieces in (1, 4, 16, 64):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step = n // pieces
        for i in range(pieces):
            d[i*step:(i+1)*step].copy_(h[i*step:(i+1)*step], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"H2D {pieces:3d} pieces: {best:.3f} ms  {n/best/1e6:.1f} GB/s")
