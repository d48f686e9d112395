"""P2P bandwidth between GPU 0 and 1 with cudaMemcpyPeer (torch copy_), one direction and both at once."""
import time, torch
n = 1 << 28  # 1 GiB of int32
a0 = torch.empty(n, dtype=torch.int32, device="cuda:0"); b0 = torch.empty(n, dtype=torch.int32, device="cuda:0")
a1 = torch.empty(n, dtype=torch.int32, device="cuda:1"); b1 = torch.empty(n, dtype=torch.int32, device="cuda:1")
s0 = torch.cuda.Stream(device="cuda:0"); s1 = torch.cuda.Stream(device="cuda:1")
def sync():
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
for _ in range(2):
    a0.copy_(a1); b1.copy_(b0)
sync()
t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s0):
        a0.copy_(a1, non_blocking=True)
sync()
dt = (time.perf_counter() - t) / 5
print("GPU1 -> GPU0 (pull on 0): %.1f GB/s" % (4 * n / dt / 1e9))
t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s0):
        a0.copy_(a1, non_blocking=True)
    with torch.cuda.stream(s1):
        b1.copy_(b0, non_blocking=True)
sync()
dt = (time.perf_counter() - t) / 5
print("both directions at once: %.1f GB/s each" % (4 * n / dt / 1e9))
