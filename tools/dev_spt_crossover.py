"""Developer probe: scalar (SPT=1) vs packed (SPT=2) kernels over K and H, device-resident step time (run once per SPT)."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from dev_time_mppi import run
from mpc_rs_b200 import models
for H, dt in ((100, 0.008), (40, 0.02), (16, 0.05)):
    for K in (16384, 32768, 49152, 65536, 75776):
        run(models.NL, H, K, dt, "f32", reps=30)
