"""``from high_order_layers_torch.layers import *`` (networks.py:7,24): high_order_fc_layers,
MaxAbsNormalization, the layer classes -- and, like the real module, ``torch``, ``math``, ``Tensor``, ``Any``
leak through the star import (Quirk Q7: networks.py uses them without importing them)."""
import math  # noqa: F401
from typing import Any, Optional  # noqa: F401

import torch  # noqa: F401
from torch import Tensor, nn  # noqa: F401

from hoir_b200.layers import (FourierSeries, HighOrderLayer, MaxAbsNormalization,  # noqa: F401
                              PiecewiseDiscontinuousPolynomial, PiecewisePolynomial, Polynomial,
                              high_order_fc_layers)
