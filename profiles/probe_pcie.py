"""Host<->device copy bandwidth probe for the e2e discussion.

    python profiles/probe_pcie.py                 one GPU: pinned and pageable, both directions
    python profiles/probe_pcie.py --concurrent    1, 2, 4, .. all GPUs of the box copying device -> pinned
                                                  host AT THE SAME TIME (one process per GPU): the box's
                                                  aggregate D2H ceiling, which bounds the multi-GPU e2e number
Writes the concurrent result as JSON on stdout ({"1": GB/s per GPU, "2": .., "aggregate": {...}, "topology": ...}).
"""
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time


def worker(dev, n_bytes, start_evt, q, bind):
    import torch
    torch.cuda.set_device(dev)
    if bind:
        try:
            pr = torch.cuda.get_device_properties(dev)
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            if node >= 0:
                cpus = set()
                for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    h.copy_(d); torch.cuda.synchronize()
    q.put(("ready", dev))
    start_evt.wait()
    reps = 6
    t = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    q.put(("done", dev, reps * n_bytes / dt / 1e9))


def concurrent(n_bytes=1 << 30):
    import torch
    total = torch.cuda.device_count()
    out = {"bytes_per_copy": n_bytes, "per_gpu": {}, "aggregate": {}}
    ns = [n for n in (1, 2, 4, 8) if n <= total]
    ctx = mp.get_context("spawn")
    for n in ns:
        q, evt = ctx.Queue(), ctx.Event()
        procs = [ctx.Process(target=worker, args=(d, n_bytes, evt, q, True)) for d in range(n)]
        [p.start() for p in procs]
        for _ in range(n):
            q.get()
        evt.set()
        rates = [q.get()[2] for _ in range(n)]
        [p.join() for p in procs]
        out["per_gpu"][str(n)] = min(rates)
        out["aggregate"][str(n)] = sum(rates)
        out[str(n)] = min(rates)
    try:
        out["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        out["topology"] = None
    try:
        out["numa_nodes"] = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
    except Exception:
        out["numa_nodes"] = None
    out["host_cpus"] = os.cpu_count()
    print(json.dumps(out))


def single():
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


if __name__ == "__main__":
    if "--concurrent" in sys.argv:
        concurrent()
    else:
        single()
