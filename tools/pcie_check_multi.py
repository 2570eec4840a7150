"""The host <-> device ceiling of a whole box: N processes (one per GPU, each pinned to its
GPU's NUMA node like bench.py's ranks) run concurrent pinned H2D + D2H copies of the bench's
per-step byte counts at the same time; prints per-GPU and aggregate GB/s.

    python tools/pcie_check_multi.py [ngpus]         (parent)
"""
import json, os, subprocess, sys, time

CHILD = r'''
import os, sys, time, json
import torch
idx = int(sys.argv[1]); t_start = float(sys.argv[2])
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
    pinned = True
except Exception:
    pinned = False
torch.cuda.set_device(idx)
n_in, n_out = 3_240_000_000, 3_200_000_000
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(chunks, h2d=True, d2h=True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ci, co = n_in // chunks, n_out // chunks
    for k in range(chunks):
        if h2d:
            with torch.cuda.stream(s1):
                din[k * ci:(k + 1) * ci].copy_(hin[k * ci:(k + 1) * ci], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout[k * co:(k + 1) * co].copy_(dout[k * co:(k + 1) * co], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
run(10)
while time.time() < t_start:      # all processes start their timed copies together
    time.sleep(0.001)
res = {"gpu": idx, "numa_pinned": pinned}
t = min(run(100) for _ in range(3)); res["duplex_h2d_gbs"] = n_in / t / 1e9; res["duplex_d2h_gbs"] = n_out / t / 1e9
t = min(run(100, d2h=False) for _ in range(2)); res["h2d_alone_gbs"] = n_in / t / 1e9
t = min(run(100, h2d=False) for _ in range(2)); res["d2h_alone_gbs"] = n_out / t / 1e9
print(json.dumps(res))
'''

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    t_start = time.time() + 25.0 + 2.0 * n     # allocation of 6.4 GB of pinned memory per process takes a while
    procs = [subprocess.Popen([sys.executable, "-c", CHILD, str(i), str(t_start)], stdout=subprocess.PIPE, text=True)
             for i in range(n)]
    rows = []
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("{"):
                rows.append(json.loads(line))
    agg = {k: sum(r[k] for r in rows) for k in ("duplex_h2d_gbs", "duplex_d2h_gbs", "h2d_alone_gbs", "d2h_alone_gbs")}
    print(json.dumps({"ngpus": n, "processes_reporting": len(rows), "aggregate": agg,
                      "aggregate_duplex_total_gbs": agg["duplex_h2d_gbs"] + agg["duplex_d2h_gbs"], "per_gpu": rows,
                      "note": "all processes copy at the same time (the per-process phases are not re-synchronised: "
                              "alone-figures of different GPUs may overlap partially)"}))

if __name__ == "__main__":
    main()
