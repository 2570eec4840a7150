"""PCIe sanity: concurrent pinned H2D + D2H of the bench's per-step byte counts."""
import time, torch
n_in, n_out = 3_240_000_000, 3_200_000_000
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(chunks):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ci, co = n_in // chunks, n_out // chunks
    for k in range(chunks):
        with torch.cuda.stream(s1):
            din[k * ci:(k + 1) * ci].copy_(hin[k * ci:(k + 1) * ci], non_blocking=True)
        with torch.cuda.stream(s2):
            hout[k * co:(k + 1) * co].copy_(dout[k * co:(k + 1) * co], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for chunks in (1, 100, 100):
    t = run(chunks)
    print(f"chunks {chunks}: {t*1e3:.1f} ms  H2D+D2H concurrent: {n_in/t/1e9:.1f} + {n_out/t/1e9:.1f} GB/s")
torch.cuda.synchronize(); t0 = time.perf_counter(); din.copy_(hin, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter() - t0; print(f"H2D alone {n_in/t/1e9:.1f} GB/s")
t0 = time.perf_counter(); hout.copy_(dout, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter() - t0; print(f"D2H alone {n_out/t/1e9:.1f} GB/s")
