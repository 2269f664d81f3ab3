"""Developer probe: MPCB_F64 (reference order) against MPCB_F64_FAST (folded formulas, FMA) — device-resident step time on
the config #4 MPPI shape (4096 controllers x 8192 samples x H = 8, model NL6), the shipped NL6 shape and configs[1]."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from dev_time_mppi import run
from mpc_rs_b200 import models

if __name__ == "__main__":
    for prec in ("f32", "f64", "f64fast"):
        run(models.NL6, 8, 8192, 0.15, prec, C_=4096, reps=5)
    for prec in ("f64", "f64fast"):
        run(models.NL6, 8, 800000, 0.15, prec, reps=20)
        run(models.NL, 100, 65536, 0.008, prec, reps=20)
