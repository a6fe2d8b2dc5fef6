"""Fused FIC objective + gradient (obj_fun_norm + dlogp_dcov_par) against the literal oracle: rel <= 1e-8."""
import numpy as np
import pytest

from oracle import reduced_model as red
from oracle import ref_model as rm
from tests import cases
from tests.test_vi_gpu import _check, _check_ill_conditioned, _names, _run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["config1", "config2", "config3", "config5"])
def test_fic_matches_literal_oracle(ctx, case):
    c = {"config1": lambda: cases.config1(), "config2": lambda: cases.config2(),
         "config3": lambda: cases.config3(n=2000, m=200), "config5": lambda: cases.config5(n=3000, m=300)}[case]()
    obj, grad = _run(ctx, c, model="fic")
    if case == "config3":
        return _check_ill_conditioned("fic", c, obj, grad)
    obj_ref, g_ref = rm.fic_obj_grad(c["cov_par"], c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], c["delta"])
    _check(obj, grad, obj_ref, g_ref, _names(c["cov_par"]))


def test_fic_headline_knots_against_reduced_oracle(ctx):
    c = cases.config5(n=12000, m=1024)
    cp = c["cov_par"]
    obj, grad = _run(ctx, c, model="fic")
    obj_ref, g_ref = red.fic_obj_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
    _check(obj, grad, obj_ref, g_ref, _names(cp))


def test_fic_objective_only_and_ragged(ctx):
    c = cases.config5(n=9001, m=77, seed=5)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], c["mu"])
    obj, g = ctx.gauss_obj_grad("fic", "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"], want_grad=False)
    S12, S22, _ = rm.assemble(cp, "ard", c["x"], c["xu"], c["delta"])
    ref = rm.obj_fun_norm(c["mu"], rm.fic_Z(cp, S12, S22, c["delta"]), S12, S22, c["y"])
    assert g is None and obj == pytest.approx(ref, rel=1e-8)
