import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
h = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: h.copy_(x, non_blocking=True)), ("h2d", lambda: x.copy_(h, non_blocking=True))):
    for _ in range(2): fn(); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, 5*(1<<30)/(time.perf_counter()-t)/1e9, "GB/s")
# two streams, halves
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = 1<<29
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h[:half].copy_(x[:half], non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(x[half:], non_blocking=True)
torch.cuda.synchronize()
print("d2h 2 streams", 5*(1<<30)/(time.perf_counter()-t)/1e9, "GB/s")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("lscpu | head -20; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)
