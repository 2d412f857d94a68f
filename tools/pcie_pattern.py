"""D2H copy time inside a queue of kernels (host far ahead of the GPU), copy engine vs a copy kernel into mapped pinned memory."""
import torch, time
n = 2 * 1024 * 1024 + 128 * 1024
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
x = torch.randn(4096, 4096, device="cuda")
def busy(ms):
    for _ in range(int(ms * 4)):
        y = x @ x
torch.cuda.synchronize()
for mode in ("copy engine, nothing queued behind", "copy engine, kernels queued behind", "synced before copy, kernels behind"):
    ts = []
    evs = []
    for rep in range(8):
        busy(3)
        if mode.startswith("synced"): torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); h1.copy_(d1, non_blocking=True); h2.copy_(d2, non_blocking=True); b.record()
        if "kernels" in mode and "nothing" not in mode: busy(3)
        evs.append((a, b))
        if "nothing" in mode: torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"{mode:45s}", " ".join(f"{a.elapsed_time(b)*1e3:7.1f}" for a, b in evs), "us")
