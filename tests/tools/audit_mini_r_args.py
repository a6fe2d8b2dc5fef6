"""Audit of oracle/mini_r against silently ignored arguments -- TEST TOOL.

Re-runs the golden generation (tests/tools/make_golden_r.py) with every call of a mini-R builtin instrumented and
prints each (builtin, named arguments) combination the reference's R code actually uses, with its count.  A builtin
that dropped a named argument it does not implement would change results silently (plogis(log.p = TRUE) once did);
this list is what has to be -- and was -- checked by hand against the builtins' implementations.

    python tests/tools/audit_mini_r_args.py
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle.mini_r import interp as RI   # noqa: E402

seen = collections.Counter()
_orig = RI.Interp.apply_function


def _patched(self, f, args, env=None):
    if isinstance(f, RI.Builtin):
        names = tuple(sorted(nm for nm, _ in args if nm))
        if names and f.name != "list":
            seen[(f.name, names)] += 1
    return _orig(self, f, args, env)


RI.Interp.apply_function = _patched

import make_golden_r as mg   # noqa: E402

mg.generate()
for (name, kws), count in sorted(seen.items()):
    print("%-16s %-40s %d" % (name, ", ".join(kws), count))
