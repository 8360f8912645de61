"""PCIe copy rates of the box: H2D / D2H alone, H2D split over two streams, H2D while a D2H stream runs (the e2e leg of bench.py does
all three at once).  usage: python tools/h2d_bw.py"""
import subprocess
import torch


def rate(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return reps * nbytes / e0.elapsed_time(e1) / 1e6


for mb in (16, 32, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    print("H2D %3d MiB: %.1f GB/s   D2H: %.1f GB/s" % (mb, rate(lambda: d.copy_(h, non_blocking=True), n), rate(lambda: h.copy_(d, non_blocking=True), n)))

n = 32 << 20
h1, h2 = (torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2))
d1, d2 = (torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()


def two_h2d():
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


def h2d_with_d2h():
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


print("2 x 32 MiB H2D on two streams: %.1f GB/s aggregate" % rate(two_h2d, 2 * n))
print("32 MiB H2D beside 32 MiB D2H: %.1f GB/s per direction" % rate(h2d_with_d2h, n))
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"], capture_output=True, text=True).stdout)
