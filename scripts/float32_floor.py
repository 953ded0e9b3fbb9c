#!/usr/bin/env python
"""How far is each GPU code path from the double-state yardstick
(tests/golden/f64_state_gradients.npz, scripts/make_f64_state_gradients.py),
next to the float32 oracle's own distance?  One subprocess per mode because
the library reads its switches once.

  python scripts/float32_floor.py            # prints one JSON line per mode
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "scripts"))
import make_f64_state_gradients as mk
from quantum_b200 import ops
z = np.load(os.path.join(%(root)r, "tests", "golden", "f64_state_gradients.npz"))
out = {"mode": %(mode)r}
for name, fn in (("c2", mk.c2), ("c4", mk.c4)):
    progs, names, vals, sums, down = fn()
    g = ops.tfq_adj_grad(progs, names, vals, sums, down).astype(np.float64)
    ref, orc32 = z[name + "_f64_state"], z[name + "_oracle_f32"].astype(np.float64)
    out[name] = {"gpu_vs_f64": float(np.abs(g - ref).max()),
                 "oracle_vs_f64": float(np.abs(orc32 - ref).max()),
                 "gpu_vs_oracle": float(np.abs(g - orc32).max()),
                 "scale": float(np.abs(ref).max())}
out["jit_launches"] = ops.get_context().profile_read()["jit_pass_launches"]
print(json.dumps(out))
'''

MODES = {
    "interpreter": {"TFQB_JIT": "0"},
    "specialised": {"TFQB_JIT_MIN_AMPS": "0"},
    "specialised, exact gates (TFQB_JIT_PHASE_FREE=0)":
        {"TFQB_JIT_MIN_AMPS": "0", "TFQB_JIT_PHASE_FREE": "0"},
    "specialised, fp64 warp shuffles (TFQB_GRAD_SHUFFLE=double)":
        {"TFQB_JIT_MIN_AMPS": "0", "TFQB_GRAD_SHUFFLE": "double"},
    "specialised, no diagonal runs (TFQB_JIT_NO_DIAG_RUN=1)":
        {"TFQB_JIT_MIN_AMPS": "0", "TFQB_JIT_NO_DIAG_RUN": "1"},
    "specialised, no real / x-imag adjoint steps (TFQB_JIT_NO_ADJ_REAL=1)":
        {"TFQB_JIT_MIN_AMPS": "0", "TFQB_JIT_NO_ADJ_REAL": "1"},
}

if __name__ == "__main__":
    for mode, env in MODES.items():
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "mode": mode}],
                           capture_output=True, text=True, env=e, timeout=1800)
        if r.returncode != 0:
            print(json.dumps({"mode": mode, "error": r.stderr[-800:]}))
        else:
            print(r.stdout.strip().splitlines()[-1], flush=True)
