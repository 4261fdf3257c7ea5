import torch, time
dev = torch.device("cuda", 0)
n = 200_000_000
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device=dev)
h2 = torch.empty(50_000_000, dtype=torch.float64).pin_memory()
d2 = torch.empty(50_000_000, dtype=torch.float64, device=dev)
for _ in range(2):
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter() - t
print("H2D 1.6 GB: %.1f GB/s" % (n * 8 / t1 / 1e9))
t = time.perf_counter(); h2.copy_(d2, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter() - t
print("D2H 0.4 GB: %.1f GB/s" % (50e6 * 8 / t2 / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t3 = time.perf_counter() - t
print("both concurrently: %.1f ms (H2D alone %.1f, D2H alone %.1f)" % (t3 * 1e3, t1 * 1e3, t2 * 1e3))
# chunked H2D in 8 pieces
torch.cuda.synchronize(); t = time.perf_counter()
c = n // 8
for i in range(8): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
torch.cuda.synchronize(); print("H2D in 8 chunks: %.1f GB/s" % (n * 8 / (time.perf_counter() - t) / 1e9))
