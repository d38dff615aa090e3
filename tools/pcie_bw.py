"""Host<->device copy bandwidth of the box (pinned memory), one direction at a time and both at once: the ceiling of
the e2e leg's image read-back + command upload."""
import time
import torch

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


for _ in range(2):
    run(True, True, 2)
print("H2D alone   %.1f GB/s" % run(True, False))
print("D2H alone   %.1f GB/s" % run(False, True))
print("both        %.1f GB/s each" % run(True, True))
# small transfers like the bench's: 1.2 MB D2H + 0.9 MB H2D per frame
m = 1228800
hs = torch.empty(m, dtype=torch.uint8).pin_memory()
ds = torch.empty(m, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    hs.copy_(ds, non_blocking=True)
torch.cuda.synchronize()
print("1.2 MB D2H copies back to back: %.1f GB/s, %.1f us each" % (2000 * m / (time.perf_counter() - t0) / 1e9, (time.perf_counter() - t0) / 2000 * 1e6))
