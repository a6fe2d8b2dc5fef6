"""The C-ABI library loads and exports every symbol include/srgp.h declares (no compute calls, no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "srgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srgp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from sparsergps_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libsrgp.so does not export " + n


def test_python_signatures_cover_header():
    from sparsergps_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.load()


def test_version_and_no_device_error():
    from sparsergps_b200 import _lib
    lib = _lib.load()
    assert lib.srgp_version() == 100
    import torch
    if not torch.cuda.is_available():
        h = _lib.vp()
        st = lib.srgp_ctx_create(0, ctypes.byref(h))
        assert st == _lib.ERR_CUDA          # fails loudly: no CPU fallback
        assert b"no CPU fallback" in lib.srgp_last_error()


def test_scalar_exports_known_answers():
    # analytic known answers; host code in the library, no GPU needed
    import numpy as np
    from sparsergps_b200 import rcpp_exports as R
    assert R.cov_fun_sqrd_expC([0.0], [1.0], {"sigma": 1, "l": 1}) == pytest.approx(0.6065306597126334, rel=1e-15)
    cp = {"sigma": 2, "l1": 1, "l2": 2}
    assert R.cov_fun_sqrd_exp_ardC([1.0, 2.0], [0.0, 0.0], cp, ["l1", "l2"]) == pytest.approx(1.4715177646857693, rel=1e-15)
    assert R.dsqexp_dsigmaC([0.3], [0.3], {"sigma": 2, "l": 1})["derivative"] == pytest.approx(8.0)
    assert R.dsqexp_dtauC([0.3, 1], [0.3, 1], {"tau": 0.5})["derivative"] == pytest.approx(0.5)
    assert R.dsqexp_dtauC([0.3, 1], [0.3, 2], {"tau": 0.5})["derivative"] == 0.0
    assert R.real_to_bounded([0.0], [3.0], [1.0])[0] == pytest.approx(2.0)
    np.testing.assert_allclose(R.real_to_pos(R.pos_to_real([0.7, 2.0])), [0.7, 2.0], rtol=1e-15)
    # exp kernel: L1 distance in the covariance, L2 in the derivatives (quirk Q8)
    assert R.cov_fun_expC([0.0, 0.0], [3.0, 4.0], {"sigma": 1, "l": 1}) == pytest.approx(np.exp(-7.0))
    assert R.dexp_dsigmaC([0.0, 0.0], [3.0, 4.0], {"sigma": 1, "l": 1})["derivative"] == pytest.approx(2 * np.exp(-5.0))


def test_library_sass_is_blackwell_native():
    """The built library must contain the sm_100a tensor-core path, not a recompiled legacy one: tcgen05.mma (SASS
    UTCIMMA) with TMEM loads (LDTM) and TMA bulk copies (UBLKCP) in the INT8 row-pass kernels."""
    import re
    import shutil
    import subprocess
    from sparsergps_b200 import build
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", build.LIB], capture_output=True, text=True).stdout
    cur, per = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = {"UTCIMMA": 0, "LDTM": 0, "UBLKCP": 0}
        elif cur:
            for k in per[cur]:
                if k in line:
                    per[cur][k] += 1
    gram = [v for f, v in per.items() if "i8_gram2_kernel" in f]
    km = [v for f, v in per.items() if "i8_km2_kernel" in f]
    assert gram and km, "INT8 tensor-core kernels missing from libsrgp.so"
    for v in gram + km:
        assert v["UTCIMMA"] == 28 and v["LDTM"] >= 1 and v["UBLKCP"] >= 1, v      # 28 slice pairs per k-step (7 slices)
