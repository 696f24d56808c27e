"""ddmi_b200: B200-native D2C-VAE continuous decoding (the INR query of mlvlab/DDMI).

Drop-in for the reference's decoder modules (models/d2c_vae/mlp.py) and NeRF
renderer (utils/nerf_helpers.py); hand-written sm_100a CUDA behind a C ABI
(include/ddmi_b200.h).  No CPU path, no eager fallback.
"""
from .mlp import MLP, MLP3D, MLPVideo, MLPNeRF  # noqa: F401
from .general_utils import (convert_to_coord_format_2d, convert_to_coord_format_3d,  # noqa: F401
                            get_scale_injection, make_3d_grid)
from . import nerf_helpers  # noqa: F401
from . import generation  # noqa: F401
from .plane_tail import PlaneTail  # noqa: F401

__all__ = ['MLP', 'MLP3D', 'MLPVideo', 'MLPNeRF', 'convert_to_coord_format_2d',
           'convert_to_coord_format_3d', 'get_scale_injection', 'make_3d_grid', 'nerf_helpers', 'generation', 'PlaneTail']
