"""Write-only / copy / read-only HBM throughput with plain torch kernels (what ceiling a store-dominated kernel can expect).
Usage: python tools/hbm_write_probe.py [GiB]"""
import sys
import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gib * (1 << 30) / 4)
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


nbytes = n * 4
for name, fn, moved in (("fill_ (write only)", lambda: a.fill_(1.0), nbytes), ("zero_ (memset)", lambda: a.zero_(), nbytes),
                        ("copy_ (1 read : 1 write)", lambda: b.copy_(a), 2 * nbytes), ("sum (read only)", lambda: a.sum(), nbytes),
                        ("mul_ in place (1:1 same lines)", lambda: a.mul_(1.0001), 2 * nbytes)):
    ms = timed(fn)
    print(f"{name:36s} {ms:8.3f} ms  {moved / ms / 1e6:8.1f} GB/s")
