"""Host<->device copy bandwidth probe (pinned and pageable) for the e2e discussion."""
import time
import torch
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hg = torch.empty(n, dtype=torch.uint8)
for name, h in (("pinned", hp), ("pageable", hg)):
    for direction in ("d2h", "h2d"):
        torch.cuda.synchronize()
        best = 0
        for _ in range(3):
            t = time.perf_counter()
            if direction == "d2h":
                h.copy_(d, non_blocking=False)
            else:
                d.copy_(h, non_blocking=False)
            torch.cuda.synchronize()
            best = max(best, n / (time.perf_counter() - t) / 1e9)
        print(f"{name} {direction}: {best:.1f} GB/s")
