"""Host-to-device bandwidth of one 57.7 MB batch of raw trials from pinned memory: 1 / 2 / 4 concurrent copy streams."""
import time
import torch
n = 256 * 128 * 440
h = [torch.randn(n).pin_memory() for _ in range(2)]
d = [torch.empty(n, device="cuda") for _ in range(2)]
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    part = n // k
    def copy(i):
        for j, st in enumerate(streams):
            with torch.cuda.stream(st):
                d[i % 2][j * part:(j + 1) * part].copy_(h[i % 2][j * part:(j + 1) * part], non_blocking=True)
    for i in range(3): copy(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 40
    for i in range(reps): copy(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{k} stream(s): {dt * 1e3:.3f} ms per batch  {n * 4 / dt / 1e9:.1f} GB/s")
