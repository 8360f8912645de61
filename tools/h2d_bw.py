import torch, time
for mb in (16, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D %d MiB: %.1f GB/s" % (mb, 10 * n / e0.elapsed_time(e1) / 1e6))
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("D2H %d MiB: %.1f GB/s" % (mb, 10 * n / e0.elapsed_time(e1) / 1e6))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"], capture_output=True, text=True).stdout)
