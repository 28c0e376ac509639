"""hoir_b200 -- B200-native (sm_100a) drop-in for the high-order layer path of
jloveric/high-order-implicit-representation.  Importable as ``hoir_b200`` (see the shim
``hoir_b200.py`` at the repository root; the directory name carries hyphens)."""
from . import _lib, config, datasets, parallel, random_sample, spec  # noqa: F401
from ._lib import HoirUnsupported, build  # noqa: F401
from .config import (Config, apply_overrides, generative_config, images_config, neighborhood_config,  # noqa: F401
                     random_interpolation_config, to_cfg)
from .datasets import image_neighborhood_dataset, image_to_dataset, neighborhood_gather  # noqa: F401
from .engine import FusedTrainer, GraphedStep, HostBatchStepper  # noqa: F401
from .rendering import generate_sequence, neighborhood_sample_generator, render_image  # noqa: F401
from .layers import (FourierSeries, HighOrderLayer, MaxAbsNormalization,  # noqa: F401
                     PiecewiseDiscontinuousPolynomial, PiecewisePolynomial, Polynomial,
                     high_order_fc_layers, segment_index)
from .networks import (Adam, GenerativeNetwork, GenNet, HighOrderMLP, Net,  # noqa: F401
                       initialize_network_polynomial_layers, node_positions)
from .random_sample import (RadialRandomImageSampleDataset, RandomImageSampleDataset,  # noqa: F401
                            generate_sample, generate_sample_radial, indices_from_grid,
                            random_radial_samples_from_image, random_symmetric_sample)
from .sparse_optimizers import SparseLion  # noqa: F401
from .trainer import ImageTrainer, NeighborhoodTrainer, RandomInterpolationTrainer  # noqa: F401
from .utils import positions_from_mesh  # noqa: F401

__version__ = "0.1.0"
