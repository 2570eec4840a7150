"""The C-ABI library: every symbol include/gppd.h declares is exported, the
host-only helpers work without a GPU, and GPU entry points fail loudly (no CPU
fallback) when no B200 is visible."""
import ctypes as C
import os
import sys
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gppd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gppd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(gp):
    L = gp._lib.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/gppd.h but not exported"


def test_host_helpers(gp, ora):
    L = gp._lib.lib()
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gppd.h")).read()
    assert L.gppd_version() == int(re.search(r"#define GPPD_VERSION (\d+)", hdr).group(1)) == 122
    for side in (0, 16):
        for tel in range(1, 5):
            for dio in range(1, 6):
                assert gp.idx(side, tel, dio) == ora.idx(side, tel, dio)
    with pytest.raises(ValueError):
        gp.idx(3, 1, 1)
    p = np.empty(8)
    assert L.gppd_phirange(p.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert p.tobytes() == ora.phirange().tobytes()
    assert L.gppd_num_windows(1000, 0) == 1 and L.gppd_num_windows(1000, 300) == 4
    # --window arithmetic of src/GPPupilDemodulation.jl:192
    tu = (2000 * np.arange(10)).astype(np.int32)
    assert gp.table_windows(tu, 59949.0, 100.0)[0] == 50004
    assert gp.table_windows(tu, 59949.0, 1.0)[0] == 500
    assert gp.table_windows(tu, 59949.0, None) == (10, 1)
    assert L.gppd_strerror(3).decode().startswith("no sm_100")


def test_options_struct_layout(gp):
    # must match `struct gppd_options` of include/gppd.h
    assert C.sizeof(gp._lib.Options) == 4 * 4 + 2 * 8 + 8 + 8 + 2 * 4
    assert gp._lib.Options.group_mask.offset == 48
    assert gp._lib.Options.xinit.offset == 16 and gp._lib.Options.rhobeg.offset == 32


def test_no_cpu_fallback(gp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the loud-failure path is for GPU-less hosts")
    with pytest.raises(gp.GppdError) as e:
        gp.Handle(0)
    assert e.value.status == 3
    t = np.arange(10.0)
    with pytest.raises(gp.GppdError):
        gp.demodulateall(t, np.zeros((10, 40), complex))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gppupildemodulation.jl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                assert "#include \"../oracle" not in txt and "oracle/" not in txt.replace("(oracle/newuoa.c)", ""), f


def test_option_flags_agree_everywhere():
    """The option bits of include/gppd.h, of the Python mirror and of the Julia binding."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gppd.h")).read()
    jl = open(os.path.join(root, "julia", "GPPDB200.jl")).read()
    sys.path.insert(0, root)
    from gppd_b200 import _lib
    flags = dict(re.findall(r"#define (GPPD_[A-Z0-9_]+) (\d+)u", hdr))
    assert {"GPPD_ONLYHIGH", "GPPD_FITOFFSETS", "GPPD_NO_RECENTER", "GPPD_KEEPRAW", "GPPD_BIG_ENDIAN",
            "GPPD_CENTER_EMPIRICAL", "GPPD_FP32"} <= set(flags)
    vals = sorted(int(v) for v in flags.values())
    assert vals == [1, 2, 4, 8, 16, 32, 64]
    for name, v in flags.items():
        assert getattr(_lib, name[5:]) == int(v), name
        m = re.search(r"const %s\s*=\s*UInt32\((\d+)\)" % name, jl)
        if m:                                   # (the binding declares the flags it uses)
            assert int(m.group(1)) == int(v), name
    assert re.search(r"const GPPD_FP32\s*=\s*UInt32\(64\)", jl)
