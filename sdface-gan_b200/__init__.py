"""sdface-gan_b200: the SDF generator's per-ray-sample field evaluation + volume rendering on B200 (sm_100a).

Import through the alias module at the repo root (``import sdface_gan_b200``) or
``importlib.import_module("sdface-gan_b200")``.  See DESIGN.md for the path and INTEGRATION.md for the drop-in boundary.
"""
from . import _lib, compat_backend, distributed, ops  # noqa: F401
from .gridencoder import GridEncoder, grid_encode  # noqa: F401
from .shencoder import SHEncoder, sh_encode  # noqa: F401
from .sdf_model import (FCGenerator, FiLMSiren, Generator, LinearLayer, MappingLinear, NGPSIRENGenerator, SirenGenerator,  # noqa: F401
                        VolumeFeatureRenderer, get_encoder, register_decoder)
from .decoder import Decoder  # noqa: F401
from .graphed import GraphedGenerator  # noqa: F401
from .sdf_utils import Munch, align_volume, default_options, generate_camera_params  # noqa: F401
