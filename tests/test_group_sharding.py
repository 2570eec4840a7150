"""One demodulateall call split over GPUs (BASELINE.json config 4, SURVEY.md section 8e
(iii)): the 8 (telescope, side) groups are independent (reference src/Modulation.jl:387-390),
so rank r runs the call with gppd_options.group_mask = its groups and the host gathers the
columns.  Here the ranks are played one after the other on one GPU: the gathered result
must be bit-identical to the unsharded call made with the same harmonic kernel (and the
unsharded call is held to the oracle)."""
import ctypes as C

import numpy as np
import pytest

import fitref
from conftest import make_case

pytestmark = pytest.mark.gpu


# The harmonic sums of complex128 arrays have two kernels: int8 tensor cores (one block serves all
# 8 groups) and FP64 DMMA (one block per group).  A call whose mask holds fewer than 5 groups takes
# the per-group kernel (less work); the sums of the two differ in the last bits (2e-14), so the
# bit-for-bit statement is made with the kernel fixed (GPPD_TENSOR_MIN_GROUPS = 1: always tensor,
# 9: never), and across kernels the fits are held to the fork envelope.
KERNELS = {"tensor": "1", "dmma": "9"}


@pytest.mark.parametrize("kernel", ["tensor", "dmma"])
@pytest.mark.parametrize("faint", [False, True])
def test_group_mask_equals_unsharded(gp, ora, monkeypatch, faint, kernel):
    n = 20_000
    tab = make_case(gp.synthetic, n, k=51, faint=faint, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    state = tab["state"]
    monkeypatch.setenv("GPPD_TENSOR_MIN_GROUPS", KERNELS[kernel])
    full = gp.demodulateall(t, z, faintparam=state, raw=True, return_info=True)
    oo, op, ol, onf = ora.demodulateall(t, z, faintparam=state, nthreads=8, return_nfev=True)
    coincide, stats = fitref.compare_fits(full[1], full[2], op, ol, full[3][:, 0], onf)
    assert coincide.sum() >= fitref.MIN_COINCIDE
    sh = gp.sharding
    for world in (2, 4, 8):
        parts, masks = [], []
        for r in range(world):
            m = sh.partition_groups(r, world)
            zz = z.copy()         # a rank must not need the other ranks' channels: poison them
            zz[:, [c for c in range(40) if c not in sh.mask_channels(m)]] = np.nan
            parts.append(gp.demodulateall(t, zz, faintparam=state, raw=True, groups=m))
            masks.append(m)
        mo, mp_, ml = sh.gather_groups(parts, masks)
        assert mo.tobytes() == np.ascontiguousarray(full[0]).tobytes() or np.array_equal(mo, full[0])
        assert mp_.tobytes() == full[1].tobytes() and ml.tobytes() == full[2].tobytes()


def test_group_mask_default_kernel_choice(gp, ora, monkeypatch):
    """The default: the unsharded call on the tensor cores, the sharded ones (4, 2, 1 groups per
    rank) on the per-group kernel: the same bits whatever the number of ranks, and against the
    unsharded call the same trajectories or forks inside the envelope (fitref)."""
    monkeypatch.delenv("GPPD_TENSOR_MIN_GROUPS", raising=False)
    n = 20_000
    tab = make_case(gp.synthetic, n, k=54, faint=True, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    state = tab["state"]
    full = gp.demodulateall(t, z, faintparam=state, raw=True, return_info=True)
    sh = gp.sharding
    for world in (2, 4, 8):
        parts, masks = [], []
        for r in range(world):
            m = sh.partition_groups(r, world)
            parts.append(gp.demodulateall(t, z, faintparam=state, raw=True, return_info=True, groups=m))
            masks.append(m)
        mo, mp_, ml = sh.gather_groups([p[:3] for p in parts], masks)
        if world == 2:
            first = (mo, mp_, ml)
        else:
            assert mp_.tobytes() == first[1].tobytes() and np.array_equal(mo, first[0])
        coincide, stats = fitref.compare_fits(mp_, ml, full[1], full[2])
        assert coincide.sum() >= fitref.MIN_COINCIDE, stats


@pytest.mark.parametrize("kernel", ["tensor", "dmma"])
def test_group_mask_windows_and_untouched_entries(gp, ora, monkeypatch, kernel):
    """Window mode with a mask; entries of the other groups are left untouched."""
    monkeypatch.setenv("GPPD_TENSOR_MIN_GROUPS", KERNELS[kernel])
    n, w = 6000, 1500
    tab = make_case(gp.synthetic, n, k=52)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    full = gp.demodulateall(t, z, raw=True, nwindow=w)
    m = 0b00100100
    o, p, l = gp.demodulateall(t, z, raw=True, nwindow=w, groups=m)
    ch = gp.sharding.mask_channels(m)
    assert np.array_equal(o[:, ch], full[0][:, ch])
    fits = [k * 32 + d for k in range(4) for d in ch if d < 32]
    assert p[fits].tobytes() == full[1][fits].tobytes() and l[fits].tobytes() == full[2][fits].tobytes()
    parts = [(o, p, l), gp.demodulateall(t, z, raw=True, nwindow=w, groups=0xff ^ m)]
    mo, mp_, ml = gp.sharding.gather_groups(parts, [m, 0xff ^ m], nwin=4)
    assert np.array_equal(mo, full[0]) and mp_.tobytes() == full[1].tobytes() and ml.tobytes() == full[2].tobytes()


def test_partial_mask_rejected_on_table_entry(gp):
    tab = make_case(gp.synthetic, 500, k=53)
    L, h = gp._lib.lib(), gp.default_handle()
    o = gp.api._options(groups=0x0f)
    vout = np.empty((500, 80), np.float32)
    par, chi2 = np.empty((32, 6)), np.empty(32)
    off = gp.synthetic.stefan_centres()
    rc = L.gppd_process_table_f32(
        h.raw, 500, tab["time_us"].ctypes.data_as(gp._lib._i32p), tab["mjd"],
        tab["volt"].ctypes.data_as(gp._lib._fp), off.view(np.float64).ctypes.data_as(gp._lib._dp),
        None, 0, None, 0, 0.0, C.byref(o), vout.ctypes.data_as(gp._lib._fp),
        par.ctypes.data_as(gp._lib._dp), chi2.ctypes.data_as(gp._lib._dp), None, None)
    assert rc == 5      # GPPD_ERR_UNSUPPORTED
    L.gppd_wait(h.raw, 0)


def test_device_resident_entry_equals_host_entry(gp, ora):
    """gppd_demodulate_f64_dev (device pointers, caller's stream) == gppd_demodulate_f64."""
    import torch
    n, w = 9000, 2500
    tab = make_case(gp.synthetic, n, k=54, faint=True, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    state = tab["state"]
    ref = gp.demodulateall(t, z, faintparam=state, raw=True, nwindow=w, return_info=True)
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(t).to(dev)
    d_z = torch.from_numpy(np.ascontiguousarray(z.T)).to(dev)          # [40][n] = column-major (n, 40)
    d_s = torch.from_numpy(state).to(dev)
    d_o = torch.zeros_like(d_z)
    nwin = 4
    d_p = torch.zeros((nwin * 32, 6), dtype=torch.float64, device=dev)
    d_c = torch.zeros(nwin * 32, dtype=torch.float64, device=dev)
    d_i = torch.zeros((nwin * 32, 4), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    L, h = gp._lib.lib(), gp.default_handle()
    o = gp.api._options()
    p = lambda x: C.c_void_p(x.data_ptr())
    st.wait_stream(torch.cuda.current_stream(dev))
    gp._lib.check(L.gppd_demodulate_f64_dev(h.raw, 2, C.c_void_p(st.cuda_stream), n, w, p(d_t), p(d_z),
                                            p(d_s), C.byref(o), p(d_o), p(d_p), p(d_c), p(d_i)))
    st.synchronize()
    assert np.array_equal(d_o.cpu().numpy().T, ref[0])
    assert d_p.cpu().numpy().tobytes() == ref[1].tobytes() and d_c.cpu().numpy().tobytes() == ref[2].tobytes()
    assert np.array_equal(d_i.cpu().numpy(), ref[3])
