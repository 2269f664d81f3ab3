import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from dev_time_mppi import run
from mpc_rs_b200 import models
run(models.NL, 200, 1 << 20, 0.004, "f32", reps=10)
run(models.NL, 100, 1 << 20, 0.008, "f32", reps=10)
run(models.NL, 8, 800000, 0.1, "f32")
