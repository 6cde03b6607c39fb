import torch, time
n = 2766086481
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
x = torch.empty(172800000, dtype=torch.float32, device='cuda')
xh = torch.empty(172800000, dtype=torch.float32, pin_memory=True)
for name, fn, nb in (("D2H", lambda: h.copy_(d, non_blocking=True), n), ("H2D", lambda: x.copy_(xh, non_blocking=True), 172800000*4)):
    for _ in range(2): fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, nb / dt / 1e9, "GB/s", dt * 1e3, "ms")
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): x.copy_(xh, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("both", dt * 1e3, "ms")
