"""quantit_b200 — B200-native (sm_100a) engine for QuantiT's block-sparse conserved-quantity tensor hot path.

Host-side mirror of the reference interface for the path (btensor::tensordot, block svd/truncate, two-site DMRG
H_eff·psi / Lanczos / environment updates) over the C ABI of libqtb.so (include/qtb.h). No CPU fallback.
"""
from .engine import (BTensor, CheckError, Context, CudaError, EngineRuntimeError, InvalidArgument, LogicError,  # noqa: F401
                     NoDeviceError, QtbError, compute_left_env, compute_right_env, default_context,
                     hamil2site_times_state, load_library, svd, dmrg, dmrg_options, tensordot, tensordot_host, two_sites_update,
                     contract, move_oc, coalesce, eigh, truncate, dmrg_logged,
                     EXPORTED_SYMBOLS, LIB_PATH)
from . import workloads  # noqa: F401
