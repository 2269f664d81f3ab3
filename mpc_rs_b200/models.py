"""Built-in device models — the tags passed where the reference passes fn pointers."""
from . import _abi as A
from .mppi import DeviceModel

# MPPI dynamics + cost
L = DeviceModel(A.MODEL_L, "L")          # examples/mppi4.rs:20-27,73-89
NL = DeviceModel(A.MODEL_NL, "NL")       # examples/mppi4-non-liner.rs:20-27,73-94
NL6 = DeviceModel(A.MODEL_NL6, "NL6")    # examples/mppi4-non-liner-ukf.rs:33-35,126-148
# UKF fx + hx
PEN_LIN = DeviceModel(A.MODEL_PEN_LIN, "PEN_LIN")  # examples/ukf-pen.rs:76-91
PEN_NL = DeviceModel(A.MODEL_PEN_NL, "PEN_NL")     # examples/ukf-pen2.rs:31-53
PEN6 = DeviceModel(A.MODEL_PEN6, "PEN6")           # examples/ukf-pen3.rs:35-63
NL6_UKF = DeviceModel(A.MODEL_NL6_UKF, "NL6_UKF")  # examples/mppi4-non-liner-ukf.rs:149-179
