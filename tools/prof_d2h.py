"""Raw D2H rate of a deriv-sized array with 1 / 2 / 4 concurrent copy streams (B200 box: 55.7 / 56.2 / 56.4 GB/s — PCIe, not the copy engine, bounds the host-pointer call)."""
import torch, time
n = 86016*105
d = torch.empty(n, dtype=torch.float64, device="cuda:0")
h = torch.empty(n, dtype=torch.float64).pin_memory()
def one():
    h.copy_(d, non_blocking=True)
def two(k):
    ss=[torch.cuda.Stream() for _ in range(k)]
    def f():
        c = n//k
        for i,s in enumerate(ss):
            with torch.cuda.stream(s):
                h[i*c:(i+1)*c].copy_(d[i*c:(i+1)*c], non_blocking=True)
    return f
for name, fn in (("1 stream", one), ("2 streams", two(2)), ("4 streams", two(4))):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
    print(name, f"{n*8/dt/1e9:.1f} GB/s  {dt*1e3:.3f} ms")
