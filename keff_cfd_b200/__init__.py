"""keff_cfd_b200 -- B200-native hot path of Keff-CFD (conduction solve + Keff flux integration).

Python here is plumbing only (ctypes over the C ABI in include/keffb200.h, torch for device
memory and torch.distributed); the product is keff_cfd_b200/csrc (hand-written sm_100a CUDA).
"""
from ._lib import F_BATCH_ADDITIVE, F_BATCH_GLOBAL, F_NO_MG, F_NO_TMA, F_ZERO_INIT, KeffB200Error, Params, Result, init, init_multi, lib, shutdown  # noqa: F401
from .api import (Options, apply_operator, boundary_flux, default_params, flux_field, interpolate, precondition, read_input_file,  # noqa: F401
                  solve2d, solve2d_batch, solve3d, structure_from_csv, structure_from_image)
