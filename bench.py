#!/usr/bin/env python
"""bench.py -- diode-samples demodulated per second (BASELINE.json metric).

Workload (BASELINE.json configs[2]): a synthetic GRAVITY+ night of 100 METROLOGY
tables of 1e5 rows each (70 bright / 30 FAINT, whole-file fits, `--center stefan`)
per GPU; one *step* = one pass of the hot path over the whole night.  Weak
scaling: every rank demodulates its own night (files are independent, no
collective on the data path).

  value : whole-job diode-samples/s with all inputs resident in HBM
          (C ABI gppd_process_table_f32_dev on the bench's CUDA streams, timed
          with CUDA events on those streams, max over ranks)
  e2e   : same metric through gppd_submit_table_f32 / gppd_wait with HOST
          (pinned) buffers: H2D of TIME+VOLT and D2H of VOLT+params inside the
          timed region
  roofline     : the dominant pass (by CUDA-event time inside the timed region)
                 against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline : the CPU oracle (restated reference algorithm, 8 threads like the
                 reference's Threads.@threads over the 8 diode groups) on a
                 bounded sample of the same night, on this box's host cores

`--impl reference` times that CPU restatement alone (the reference is Julia and
cannot run in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ROW = 644.0          # TIME 4 B + VOLT 320 B in, VOLT 320 B out
DIODES = 32
METRIC = "diode-samples demodulated/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gppd", choices=["gppd", "reference"])
    ap.add_argument("--files", type=int, default=100)
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--streams", type=int, default=5)
    ap.add_argument("--cpu-files", type=int, default=12, help="files in the CPU-baseline sample")
    ap.add_argument("--window-rows", type=int, default=0,
                    help="fit windows of this many rows instead of whole tables (reference --window); "
                         "an exploration switch, the headline workload is whole-file")
    ap.add_argument("--chains", type=int, default=1,
                    help="exploration: split the resident night into this many concurrent launch sequences")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions, polled through NVML
    every ~2 ms (nvidia-smi's loop mode is too coarse for a region of tens of ms)."""

    def __init__(self, index):
        self.index, self.rows, self.th, self.stop_flag = index, [], None, False
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                ut = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.rows.append((sm, rs, ut))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.th.join(timeout=2)
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": reasons, "samples": len(sm),
                "how": "NVML polled every ~2 ms during the resident and end-to-end timed regions"}


def pin_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU, so that the pinned staging buffers
    of the end-to-end path are allocated on the GPU's own NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False


class quiet_stdout:
    """NCCL prints its version banner on stdout; keep stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# --------------------------------------------------------------------------
def night_plan(nfiles):
    """70 % bright, 30 % FAINT, interleaved deterministically."""
    return [(k % 10) in (3, 6, 9) for k in range(nfiles)]


def generate_night(torch, gp, dev, nfiles, nrows, rank):
    """Synthetic night on the GPU (same model as gppd_b200.synthetic.make_table):
    returns time_us [F, N] int32, volt [F, N, 80] float32, mjd list, FaintStates list."""
    centres = torch.tensor(gp.synthetic.stefan_centres().view(np.float64).reshape(40, 2),
                           device=dev)
    time_us = torch.empty((nfiles, nrows), dtype=torch.int32, device=dev)
    volt = torch.empty((nfiles, nrows, 80), dtype=torch.float32, device=dev)
    faint = night_plan(nfiles)
    mjds, fss = [], []
    base = (2000 * torch.arange(nrows, device=dev, dtype=torch.int64))
    for k in range(nfiles):
        kk = rank * nfiles + k
        mjd = 59949.0 + kk / 100.0
        g = torch.Generator(device=dev)
        g.manual_seed(20230105 + kk)
        jit = torch.randint(-1, 2, (nrows,), generator=g, device=dev)
        jit[0] = 0
        tu = (base + jit).to(torch.int32)
        time_us[k] = tu
        t = tu.to(torch.float64) * 1e-6 + 86400.0 * mjd
        pscale = torch.ones(nrows, dtype=torch.float64, device=dev)
        fs = None
        if faint[k]:
            hdr = gp.synthetic.faint_header(mjd)
            fs = gp.buildfaintparameters(hdr)
            st = torch.from_numpy(gp.buildstates(fs, t.cpu().numpy())).to(dev)
            pscale = torch.where(st == 3, 3.0, torch.where(st == 1, 0.2, 1.0)).to(torch.float64)
        wt = 6.283185 * t
        for grp in range(8):
            phi_fc = torch.cumsum(0.02 * torch.randn(nrows, generator=g, device=dev, dtype=torch.float64), 0)
            phi_fc = phi_fc + float(torch.rand(1, generator=g, device=dev)) * 6.28 - 3.14
            efc = torch.polar(torch.ones_like(phi_fc), phi_fc)
            fcch = 32 + grp
            fc = torch.complex(centres[fcch, 0], centres[fcch, 1]) + 0.3 * efc
            fc = fc + 0.002 * torch.complex(torch.randn(nrows, generator=g, device=dev, dtype=torch.float64),
                                            torch.randn(nrows, generator=g, device=dev, dtype=torch.float64))
            volt[k, :, 2 * fcch] = fc.real.float()
            volt[k, :, 2 * fcch + 1] = fc.imag.float()
            par = torch.rand(16, generator=g, device=dev, dtype=torch.float64).cpu().numpy()
            for dio in range(4):
                ch = 4 * grp + dio
                amp = 0.05 + 0.45 * par[4 * dio]
                arg_a = -np.pi + 2 * np.pi * par[4 * dio + 1]
                b = 0.3 + 2.2 * par[4 * dio + 2]
                phi = -np.pi + 2 * np.pi * par[4 * dio + 3]
                a = amp * np.exp(1j * arg_a)
                mod = torch.polar(torch.ones_like(wt), b * torch.sin(wt + phi))
                noise = 0.02 * amp * torch.complex(
                    torch.randn(nrows, generator=g, device=dev, dtype=torch.float64),
                    torch.randn(nrows, generator=g, device=dev, dtype=torch.float64))
                d = torch.complex(centres[ch, 0], centres[ch, 1]) + complex(a) * pscale * efc * mod + noise
                volt[k, :, 2 * ch] = d.real.float()
                volt[k, :, 2 * ch + 1] = d.imag.float()
        mjds.append(mjd)
        fss.append(fs)
    return time_us, volt, mjds, fss


def cpu_reference_files(oracle, files, cores):
    """The restated reference algorithm (oracle) on host copies of `files`
    = [(time_us, volt, mjd, FaintStates|None)].  Returns seconds."""
    off = None
    import gppd_b200 as gp
    off = gp.synthetic.stefan_centres()
    t0 = time.perf_counter()
    for tu, v, mjd, fs in files:
        ofs = None if fs is None else oracle.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0)
        oracle.processmetrology(tu, v, mjd, offsets=off, faintparam=ofs, nthreads=cores)
    return time.perf_counter() - t0


# --------------------------------------------------------------------------
def main():
    args = parse()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import ctypes as C
    import gppd_b200 as gp
    from gppd_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libgppd has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_pinned = pin_to_gpu_numa_node(local)
    if world > 1:
        with quiet_stdout():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()          # creates the communicator (and prints NCCL's banner) now
            torch.cuda.synchronize()
    h = gp.Handle(local)
    L = _lib.lib()
    F, N, S = args.files, args.rows, max(1, min(args.streams, h.num_slots - 1))
    faint = night_plan(F)
    time_us, volt, mjds, fss = generate_night(torch, gp, dev, F, N, rank)
    out = torch.empty_like(volt)
    W = args.window_rows if 0 < args.window_rows < N else 0
    NW = (N + W - 1) // W if W else 1          # windows (jobs) per table
    if W:
        args.no_e2e = True
    params = torch.empty((F, NW * 32, 6), dtype=torch.float64, device=dev)
    chi2 = torch.empty((F, NW * 32), dtype=torch.float64, device=dev)
    info = torch.zeros((F, NW * 32, 4), dtype=torch.int32, device=dev)
    offsets = torch.tensor(gp.synthetic.stefan_centres().view(np.float64), device=dev)
    opt = gp.api._options()
    torch.cuda.synchronize()

    def p(tensor):
        return C.c_void_p(tensor.data_ptr())

    def dptr(a):
        return None if a is None else a.ctypes.data_as(_lib._dp)

    # one batched launch sequence for the whole night (gppd_process_tables_f32_dev)
    i64 = lambda v: (C.c_int64 * F)(*v)
    vp = lambda ts: (C.c_void_p * F)(*[t.data_ptr() for t in ts])
    dpp = lambda arrs: (_lib._dp * F)(*[C.cast(None, _lib._dp) if a is None else a.ctypes.data_as(_lib._dp) for a in arrs])
    b_n = i64([N] * F)
    b_w = i64([W] * F) if W else None
    b_mjd = (C.c_double * F)(*mjds)
    b_time, b_volt, b_out = vp([time_us[k] for k in range(F)]), vp([volt[k] for k in range(F)]), vp([out[k] for k in range(F)])
    b_par, b_chi, b_info = vp([params[k] for k in range(F)]), vp([chi2[k] for k in range(F)]), vp([info[k] for k in range(F)])
    b_t1 = dpp([fs.timer1 if fs else None for fs in fss])
    b_t2 = dpp([fs.timer2 if fs else None for fs in fss])
    b_n1 = i64([fs.timer1.size if fs else 0 for fs in fss])
    b_n2 = i64([fs.timer2.size if fs else 0 for fs in fss])

    bench_stream = torch.cuda.Stream(device=dev)   # a real (non-NULL) stream: events see it

    NCH = max(1, min(args.chains, h.num_slots, F))
    chain_streams = [torch.cuda.Stream(device=dev) for _ in range(NCH)] if NCH > 1 else []

    def sub(arr, idx, typ):
        return None if arr is None else (typ * len(idx))(*[arr[i] for i in idx])

    chains = []
    for c in range(NCH):
        idx = list(range(c, F, NCH))
        chains.append(dict(n=len(idx), b_n=sub(b_n, idx, C.c_int64), b_w=sub(b_w, idx, C.c_int64),
                           b_time=sub(b_time, idx, C.c_void_p), b_mjd=sub(b_mjd, idx, C.c_double),
                           b_volt=sub(b_volt, idx, C.c_void_p), b_t1=sub(b_t1, idx, _lib._dp),
                           b_n1=sub(b_n1, idx, C.c_int64), b_t2=sub(b_t2, idx, _lib._dp),
                           b_n2=sub(b_n2, idx, C.c_int64), b_out=sub(b_out, idx, C.c_void_p),
                           b_par=sub(b_par, idx, C.c_void_p), b_chi=sub(b_chi, idx, C.c_void_p),
                           b_info=sub(b_info, idx, C.c_void_p)))

    def step_resident():
        if NCH == 1:
            _lib.check(L.gppd_process_tables_f32_dev(
                h.raw, 0, C.c_void_p(bench_stream.cuda_stream), F, b_n, b_w, b_time, b_mjd, b_volt, p(offsets),
                b_t1, b_n1, b_t2, b_n2, C.byref(opt), b_out, b_par, b_chi, b_info, None))
            return
        fork = torch.cuda.Event()
        fork.record(bench_stream)
        for c, ch in enumerate(chains):
            st = chain_streams[c]
            st.wait_event(fork)
            _lib.check(L.gppd_process_tables_f32_dev(
                h.raw, c, C.c_void_p(st.cuda_stream), ch["n"], ch["b_n"], ch["b_w"], ch["b_time"],
                ch["b_mjd"], ch["b_volt"], p(offsets), ch["b_t1"], ch["b_n1"], ch["b_t2"], ch["b_n2"],
                C.byref(opt), ch["b_out"], ch["b_par"], ch["b_chi"], ch["b_info"], None))
            ev = torch.cuda.Event()
            ev.record(st)
            bench_stream.wait_event(ev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ---------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    h.enable_timing(True)
    h.pass_times(reset=True)
    launches0 = h.launches
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(bench_stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(bench_stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    passes_overlapped = h.pass_times(reset=True)
    h.enable_timing(False)
    launches = h.launches - launches0
    # ---- per-pass kernel times: a second region of the same K steps with ONE launch
    # sequence per batch.  In the throughput region above the FAINT and the bright tables
    # run as two concurrent chains (gppd_set_split_chains, default on), whose kernels
    # share the SMs and stretch each other's event-timed durations; the roofline of a
    # kernel needs its own duration.
    split_default = os.environ.get("GPPD_SPLIT_CHAINS", "1")[:1] != "0"
    passes = passes_overlapped
    if split_default:
        h.set_split_chains(False)
        step_resident()
        barrier()
        h.enable_timing(True)
        h.pass_times(reset=True)
        for _ in range(args.steps):
            step_resident()
        barrier()
        passes = h.pass_times(reset=True)
        h.enable_timing(False)
        h.set_split_chains(True)
    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    units_per_step = F * N * DIODES * world
    value = units_per_step / (ms_per_step * 1e-3)
    nfev_mean = float(info[:, :, 0].double().mean().item())

    # ---- e2e: host buffers through the C ABI ----------------------------
    e2e = None
    if not args.no_e2e:
        h_time = torch.empty((F, N), dtype=torch.int32).pin_memory()
        h_volt = torch.empty((F, N, 80), dtype=torch.float32).pin_memory()
        h_out = torch.empty((F, N, 80), dtype=torch.float32).pin_memory()
        h_par = torch.empty((F, 32, 6), dtype=torch.float64).pin_memory()
        h_chi = torch.empty((F, 32), dtype=torch.float64).pin_memory()
        h_time.copy_(time_us); h_volt.copy_(volt)
        torch.cuda.synchronize()
        offs_h = gp.synthetic.stefan_centres()

        def hp(t, typ):
            return C.cast(C.c_void_p(t.data_ptr()), typ)

        def step_e2e():
            for k in range(F):
                fs = fss[k]
                _lib.check(L.gppd_submit_table_f32(
                    h.raw, 1 + k % S, N, hp(h_time[k], _lib._i32p), mjds[k], hp(h_volt[k], _lib._fp),
                    dptr(offs_h.view(np.float64)), dptr(fs.timer1) if fs else None,
                    fs.timer1.size if fs else 0, dptr(fs.timer2) if fs else None,
                    fs.timer2.size if fs else 0, 0.0, C.byref(opt), hp(h_out[k], _lib._fp),
                    hp(h_par[k], _lib._dp), hp(h_chi[k], _lib._dp), None, None))
            for s in range(S):
                _lib.check(L.gppd_wait(h.raw, 1 + s))

        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_val = units_per_step * args.steps / float(tm.item())
        # result check of the e2e path against the resident path (same fits)
        same = bool(torch.equal(h_par.to(dev), params))
        e2e = {"value": e2e_val, "unit": "diode-samples/s",
               "h2d_bytes_per_step": int(F * (N * 4 + N * 320)),
               "d2h_bytes_per_step": int(F * (N * 320 + 32 * 7 * 8)),
               "timing": "host wall clock around the K steps, device-synchronised, max over ranks",
               "matches_resident_path": same}

    clk = clocks.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant pass ----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dom = max((k for k in passes if passes[k][1] > 0), key=lambda k: passes[k][0])
    dom_ms, dom_n = passes[dom]
    avg_ms = dom_ms / dom_n
    alg_bytes_per_launch = ALG_BYTES_PER_ROW * N * F      # one launch covers the whole night
    achieved = alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9
    fp64_peak = h.fp64_peak_tflops()
    # algorithmic FP64 work of the harmonic pass: the contraction C[48 x 8] += E^T V,
    # i.e. per (row, diode) 24 harmonics x 4 FMA; the recurrences
    # that generate E and the producers' arithmetic are overhead, not counted
    harm_flops_per_launch = F * N * 32 * (24 * 4) * 2.0
    harm_ms, harm_n = passes["harmonics"]
    fp64 = {"kernel": "harmonics",
            "achieved_tflops": harm_flops_per_launch / (harm_ms / max(harm_n, 1) * 1e-3) / 1e12 if harm_n else None,
            "peak_tflops": fp64_peak,
            "peak_source": "FP64 micro-benchmark in this run (better of DFMA and mma.sync.m8n8k4.f64)",
            "objective_calls_per_fit": nfev_mean}
    fp64["frac"] = fp64["achieved_tflops"] / fp64_peak if (fp64_peak and fp64["achieved_tflops"]) else None
    # which kernel computed the harmonic sums: the int8 tensor-core kernel takes every dense
    # table (GPPD_HARMONICS=dmma forces the FP64 DMMA kernel, csrc/harm_tc_kernels.cu)
    tensor_mode = os.environ.get("GPPD_HARMONICS", "")[:1] != "d"
    tensor = None
    if tensor_mode and harm_n:
        # issued int8 work: per 32 rows four MMAs of M = 128, K = 32, N = 144 + 144 + 240 + 192
        ops = F * N / 32.0 * 128 * 720 * 32 * 2
        i8_peak = 2.0 * float(peaks.get("bf16_tflops", 2250.0))
        tensor = {"kernel": "harmonics", "instruction": "tcgen05.mma kind::i8 (48-bit fixed point, 6 x 6 byte digits)",
                  "achieved_tops": ops / (harm_ms / harm_n * 1e-3) / 1e12, "peak_tops": i8_peak,
                  "peak_source": "2 x the dense bf16 figure of MEASURED_PEAKS.json (int8 runs at twice the bf16 rate)"
                                 if "bf16_tflops" in peaks else "nominal 4.5 POPS (B200_PROFILING.md fallback)",
                  "fp64_equivalent_tflops": fp64["achieved_tflops"]}
        tensor["frac"] = tensor["achieved_tops"] / i8_peak
        fp64["note"] = ("harmonic sums ran on the int8 tensor cores: achieved_tflops is the FP64 work they "
                        "replace (192 flop per diode-sample), not FP64 instructions issued")
    traffic = None   # DRAM bytes per launch of the dominant pass, from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if tj["workload"] == {"tables": F, "rows": N} and not W:
            traffic = tj["bytes_per_launch"].get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "avg_launch_ms": avg_ms, "launches_timed": dom_n,
                "pass_ms_per_step": {k: v[0] / args.steps for k, v in passes.items() if v[1]},
                "pass_timing": ("second region of the same %d steps in this run with one launch sequence per "
                                "batch (gppd_set_split_chains(0)): in the throughput region the FAINT and the "
                                "bright chain run concurrently and stretch each other's kernel durations"
                                % args.steps) if split_default else "throughput region",
                "pass_ms_per_step_overlapped": {k: v[0] / args.steps for k, v in passes_overlapped.items() if v[1]},
                "fp64": fp64, "tensor": tensor,
                "harmonics_kernel": "k_harm_tc (int8 tensor cores)" if tensor_mode else "k_harm_ws (FP64 DMMA)",
                "note": ("one launch of each pass covers the whole night; the harmonic pass does 192 FP64-equivalent "
                         "flop per diode-sample against 20 B: on the FP64 units (k_harm_ws) it is bound by them; "
                         "as exact fixed-point int8 MMAs (k_harm_tc) it is bound by the producers' issue rate "
                         "(digit extraction), see DESIGN.md section 5; the demod pass is the HBM-bound one")}

    # ---- CPU baseline on a bounded sample -------------------------------
    cpu = None
    if not args.no_cpu:
        import oracle
        oracle.build()
        cores = min(8, os.cpu_count() or 1)
        nf = min(args.cpu_files, F)
        files = [(time_us[k].cpu().numpy(), volt[k].cpu().numpy(), mjds[k], fss[k]) for k in range(nf)]
        secs = cpu_reference_files(oracle, files, cores)
        cpu = {"value": nf * N * DIODES / secs, "unit": "diode-samples/s", "cores": cores,
               "kind": "port",
               "sample": "first %d tables of the night (%d rows each), whole-file fits, %.1f s" % (nf, N, secs)}

    line = {
        "metric": METRIC, "value": value, "unit": "diode-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "night of %d METROLOGY tables x %d rows per GPU (70%% bright / 30%% FAINT), "
                               "whole-file fits, --center stefan (BASELINE.json configs[2])" % (F, N),
                   "tables_per_gpu": F, "rows_per_table": N, "diodes": DIODES, "e2e_slots": S,
                   "window_rows": W or None,
                   "cache": "inputs (%.1f GB per step) larger than L2" % (F * N * 324 / 1e9),
                   "sharding": "files -> ranks, no data-path collective",
                   "rank_pinned_to_gpu_numa_node": numa_pinned},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle restatement; the
    Julia package cannot run here) on this box's host cores, same night, each
    step a bounded sample of it."""
    if rank != 0:
        return
    import gppd_b200 as gp
    import oracle
    oracle.build()
    cores = min(8, os.cpu_count() or 1)
    F, N = args.files, args.rows
    nf = max(1, min(4, F))       # tables per step
    faint = night_plan(F)
    files = []
    for k in range(nf):
        tab = gp.synthetic.make_table(N, k=k, jitter=True)
        fs = None
        if faint[k]:
            hdr = gp.synthetic.faint_header(tab["mjd"])
            fs = oracle.buildfaintparameters(hdr)
            t = oracle.make_times(tab["time_us"], tab["mjd"])
            st = oracle.buildstates(fs, t)
            tab = gp.synthetic.make_table(N, k=k, jitter=True, faint=True, state=st)
        files.append((tab["time_us"], tab["volt"], tab["mjd"], fs))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_files(oracle, files[:1], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_files(oracle, files, cores)
    dt = time.perf_counter() - t0
    value = nf * N * DIODES * args.steps / dt
    sample = "%d tables x %d rows per step (of the %d-table night), whole-file fits" % (nf, N, F)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "diode-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "night of %d METROLOGY tables x %d rows per GPU (70%% bright / 30%% FAINT), "
                               "whole-file fits, --center stefan (BASELINE.json configs[2])" % (F, N),
                   "tables_per_gpu": F, "rows_per_table": N, "diodes": DIODES},
        "cpu_baseline": {"value": value, "unit": "diode-samples/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "diode-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "restated CPU reference algorithm (oracle/): the Julia package and its NEWUOA "
                "dependency cannot be installed in this image",
    }))


if __name__ == "__main__":
    main()
