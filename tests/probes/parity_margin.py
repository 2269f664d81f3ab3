"""Developer probe: how far the FP32 path is from the f64 oracle on the parity shapes (the tests only assert bounds)."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_mppi_gpu as T

for case, K in (("L_shipped", 16384), ("NL_shipped", 16384), ("NL_h100", 16384), ("NL_h100", 65536), ("NL6_shipped", 16384)):
    errs, cerr, ok = [], [], True
    for u_g, u_o, ig, io, c_g, c_o in T.closed_loop(case, K, "f32", eps_dtype=np.float32):
        errs.append(T.rel_err(u_g, u_o))
        near = c_o > io["max"] - 50
        cerr.append(np.max(np.abs(c_g[near] - c_o[near])))
        ok = ok and ig["argmax"] == io["argmax"]
    print(f"{case:12s} K={K:6d}  u rel err per step: " + " ".join(f"{e:.2e}" for e in errs) +
          "  | max cost err near best: " + " ".join(f"{e:.1e}" for e in cerr) + f"  argmin {'ok' if ok else 'MISMATCH'}")
