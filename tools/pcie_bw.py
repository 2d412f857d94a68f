import torch, time
for mb in (0.125, 1, 4, 16, 64):
    n = int(mb * 1024 * 1024)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if direction == "d2h": h.copy_(d, non_blocking=True)
            else: d.copy_(h, non_blocking=True)
            b.record(); torch.cuda.synchronize()
        print(f"{direction} {mb} MiB: {a.elapsed_time(b)*1e3:.1f} us -> {n/a.elapsed_time(b)/1e6:.2f} GB/s")
