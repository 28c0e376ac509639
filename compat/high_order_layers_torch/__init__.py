"""Import alias: makes ``hoir_b200`` answer to the name of the dependency the reference imports
(``high-order-layers-torch ^2.4.2``, /root/reference/pyproject.toml:21; star-imported at
/root/reference/high_order_implicit_representation/networks.py:7-8,22,24 and
single_image_dataset.py:13), so the reference's own modules run on the B200 kernels WITHOUT editing their
imports: put ``<this repository>/compat`` on ``PYTHONPATH`` ahead of (or instead of) the real package.  It lives
under ``compat/`` -- NOT at the repository root -- so that a real install of the dependency is never shadowed by
accident (``scripts/verify_real_dep.py`` and ``bench.py --impl reference`` look for the real one).

Only the names the reference uses are provided; everything resolves to the CUDA-only
implementations in ``high-order-implicit-representation_b200/``.
"""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

import hoir_b200 as _impl  # noqa: E402,F401

__version__ = "2.4.2+hoir_b200"
