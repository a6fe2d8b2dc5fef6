"""K5: trace-term entry points on materialised inputs and the fused all-theta Omega o dK reduction."""
import numpy as np
import pytest

from oracle import ref_kernels as rk
from oracle import ref_model as rm
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["config1", "config2", "config3"])
def test_trace_term_fun_matches_oracle(ctx, case):
    from sparsergps_b200 import vi_functions as V
    c = getattr(cases, case)() if case != "config3" else cases.config3(n=6000, m=130)
    cp = c["cov_par"]
    S12, S22, _ = rm.assemble(cp, c["cov_fun"], c["x"], c["xu"], c["delta"])
    ref = rm.trace_term_fun(cp, S12, S22, c["delta"])
    got = V.trace_term_fun(cp, S12, S22, c["delta"], ctx=ctx)
    assert got == pytest.approx(ref, rel=1e-9)
    assert V.dtrace_term_dtau(cp, got) == pytest.approx(rm.dtrace_term_dtau(cp, ref), rel=1e-9)


def test_dtrace_term_dcov_par(ctx):
    from sparsergps_b200 import vi_functions as V
    rng = np.random.default_rng(0)
    for n in (1, 7, 1000, 300001):
        a = rng.normal(size=n)
        cp = {"sigma": 1.0, "tau": 0.7}
        assert V.dtrace_term_dcov_par(cp, a, ctx=ctx) == pytest.approx(rm.dtrace_term_dcov_par(cp, a), rel=1e-12, abs=1e-12)


@pytest.mark.parametrize("cov_fun,d", [("ard", 5), ("ard", 8), ("sqexp", 2), ("ard", 11)])
def test_omega_dk_reduce_never_materialises_dk(ctx, cov_fun, d):
    """sum(Omega * dSigma12/dtheta) for every theta equals the sums over the oracle's materialised matrices."""
    from sparsergps_b200 import vi_functions as V
    rng = np.random.default_rng(d)
    n, m = 1500, 70
    x = rng.normal(size=(n, d))
    xu = np.vstack([rng.normal(size=(m - 2, d)), x[3], x[700]])          # coincident rows -> tau term
    Om = rng.normal(size=(n, m))
    if cov_fun == "ard":
        cp = cases.ard_par(1.2, rng.uniform(0.6, 1.8, d), 0.45)
        ln = ["l%d" % (i + 1) for i in range(d)]
        ref = {k: float(np.sum(Om * rk.dsig_dtheta_ardC(x, xu, cp, "ard", k, ln))) for k in cp}
    else:
        cp = {"sigma": 1.2, "l": 0.9, "tau": 0.45}
        ref = {k: float(np.sum(Om * rk.dsig_dthetaC(x, xu, cp, "sqexp", k))) for k in cp}
    got = V.omega_dk_reduce(cp, cov_fun, x, xu, Om, ctx=ctx)
    scale = max(abs(v) for v in ref.values())
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=1e-9, abs=1e-11 * scale), k


def test_reference_shaped_gradient_calls(ctx):
    from sparsergps_b200 import vi_functions as V
    c = cases.config2(n=800, m=48)
    out = V.delbo_dcov_par(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"], ctx=ctx)
    obj_ref, g_ref = rm.vi_obj_grad(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"])
    assert out["objective"] == pytest.approx(obj_ref, rel=1e-8)
    for k in g_ref:
        assert out["gradient"][k] == pytest.approx(g_ref[k], rel=1e-8, abs=1e-10)
        assert out["trans_par"][k] == pytest.approx(np.log(c["cov_par"][k]))
    out = V.dlogp_dcov_par(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"], ctx=ctx)
    obj_ref, g_ref = rm.fic_obj_grad(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"])
    assert out["objective"] == pytest.approx(obj_ref, rel=1e-8)
    for k in g_ref:
        assert out["gradient"][k] == pytest.approx(g_ref[k], rel=1e-8, abs=1e-10)
