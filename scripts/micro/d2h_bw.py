"""D2H bandwidth of one 980 MB page-locked copy on 1 / 2 / 4 streams (B200 box of this pool: 54.7 / 57.1 / 57.0 GB/s): why the host fill sends MDK in two halves."""
import torch, time
n = 980 * 1000 * 1000 // 8
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
d.fill_(1.0); torch.cuda.synchronize()
def run(nstreams, chunks_per_stream=1):
    ss = [torch.cuda.Stream() for _ in range(nstreams)]
    parts = nstreams * chunks_per_stream
    step = (n + parts - 1) // parts
    torch.cuda.synchronize(); t = time.perf_counter()
    for rep in range(5):
        for p in range(parts):
            with torch.cuda.stream(ss[p % nstreams]):
                h[p*step:(p+1)*step].copy_(d[p*step:(p+1)*step], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(f"streams={nstreams} chunks/stream={chunks_per_stream}: {dt*1e3:.2f} ms, {n*8/dt/1e9:.1f} GB/s", flush=True)
for k in (1, 2, 4):
    run(k)
run(1, 8); run(2, 8)
# H2D
d2 = torch.empty(42_000_000 // 8, dtype=torch.float64, device="cuda"); h2 = torch.empty(42_000_000 // 8, dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); print("h2d 42MB: %.2f ms" % ((time.perf_counter() - t) / 10 * 1e3))
