"""mpc_rs_b200 — B200 (sm_100a) MPPI rollout and batched UKF hot paths of teruyamato0731/mpc-rs.

The compute lives in libmpc_b200.so (hand-written CUDA behind the C ABI of include/mpc_b200.h); this
package is the thin host-side mirror of the reference's public types (mppi::Mppi, ukf::UnscentedKalmanFilter,
ukf2::UnscentedKalmanFilter, gaussian::Gaussian).  There is no CPU fallback: without the built library
and a CUDA device every constructor raises.
"""
from . import _abi
from ._abi import MpcB200Error
from .mppi import DeviceModel, Mppi, MppiError, UserModel, check_user_source, comm_unique_id, user_model
from . import models
from . import ukf
from .ukf import BatchedUkf, UnscentedKalmanFilter, UkfError, UserUkfModel, check_user_ukf_source, user_ukf_model
from .gaussian import Gaussian

__all__ = ["Mppi", "MppiError", "DeviceModel", "UserModel", "user_model", "check_user_source", "UserUkfModel", "user_ukf_model", "check_user_ukf_source", "MpcB200Error", "models", "comm_unique_id", "ukf", "BatchedUkf",
           "UnscentedKalmanFilter", "UkfError", "Gaussian"]
