"""Generates tests/golden/golden_v1.npz (round 1 cases; `--v1` to rewrite it) and golden_v2.npz / golden_v3.npz (round 2 cases) with the CPU checker (oracle/).  The Rust reference cannot run in this image,
so these vectors are NOT reference output: they freeze the checker's results (which reproduce every count the
reference's tests pin) so that a later change to oracle/ or to the kernels is caught against a committed artefact.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ellref_py as R  # noqa: E402
from golden_cases import CASES, CASES_V2, CASES_V3, run_stepwise, run_stepwise_nonsym  # noqa: E402

FIELDS = ("xbest", "has_xbest", "num_iters", "exit_status", "gamma", "tsq", "xc", "kappa")
out2 = {}
for name, make in CASES_V2.items():
    o = R.solve(make(R))
    for f in FIELDS:
        out2[f"{name}/{f}"] = getattr(o, f)
st = run_stepwise_nonsym(lambda **kw: R.Space(0, **kw), lambda s, w, g, b0, b1: s.update(w, g, b0, b1, single=False), lambda s: s.state())
for k, v in st.items():
    out2[f"stepwise_nonsym/{k}"] = v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"), **out2)
print("wrote", len(out2), "arrays to golden_v2.npz")
out3 = {}
for name, make in CASES_V3.items():
    o = R.solve(make(R))
    for f in FIELDS:
        out3[f"{name}/{f}"] = getattr(o, f)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v3.npz"), **out3)
print("wrote", len(out3), "arrays to golden_v3.npz")
if "--v1" not in sys.argv:
    sys.exit(0)

out = {}
for name, make in CASES.items():
    o = R.solve(make(R))
    for f in ("xbest", "has_xbest", "num_iters", "exit_status", "gamma", "tsq", "xc", "kappa"):
        out[f"{name}/{f}"] = getattr(o, f)
st = run_stepwise(lambda **kw: R.Space(0, **kw), lambda s, w, g, b0, b1: s.update(w, g, b0, b1, single=False), lambda s: s.state())
for k, v in st.items():
    out[f"stepwise/{k}"] = v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
print("wrote", len(out), "arrays")
