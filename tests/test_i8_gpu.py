"""The two tensor-core engines of the row passes must tell the same story: the INT8 engine (tcgen05.mma.kind::i8 with
error-free digit splitting, the default) against the FP64 DMMA engine (SRGP_TENSOR=dmma, read once per process --
hence the subprocesses).  Also shapes that stress the INT8 path's padding: m below one tile, m not a multiple of 64,
n below one chunk, a knot that is a data row (the q == 2^62 marker of quirk Q4)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
from tests import cases
from sparsergps_b200.context import Context
out = {}
with Context(0) as ctx:
    for name, c in (("config5", cases.config5(n=30011, m=300, seed=5)), ("config2", cases.config2()),
                    ("config3_coincident", cases.config3(n=3000, m=70)), ("tiny", cases.config5(n=97, m=5, seed=6)),
                    ("config1", cases.config1())):
        cp = c["cov_par"]
        ctx.set_data(c["x"], c["y"], c["mu"])
        for model in ("vi", "fic"):
            obj, grad = ctx.gauss_obj_grad(model, c["cov_fun"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
            out[name + "/" + model] = [obj] + [float(g) for g in grad]
print("RESULT " + json.dumps(out))
''' % ROOT


def _run(engine):
    env = dict(os.environ)
    env.pop("SRGP_TENSOR", None)
    if engine:
        env["SRGP_TENSOR"] = engine
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_int8_and_dmma_engines_agree():
    i8, dm = _run(None), _run("dmma")
    assert i8.keys() == dm.keys()
    for k in i8:
        a, b = np.array(i8[k]), np.array(dm[k])
        # both engines are ~1e-13 from the exact value on these configs; config 3 (cond 4e4) amplifies to ~1e-10
        tol = 1e-8 if k.startswith("config3") else 1e-10
        assert a[0] == pytest.approx(b[0], rel=tol), k
        np.testing.assert_allclose(a[1:], b[1:], rtol=tol, atol=tol * np.max(np.abs(b[1:])), err_msg=k)
