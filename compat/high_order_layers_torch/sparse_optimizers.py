"""``from high_order_layers_torch.sparse_optimizers import SparseLion`` (networks.py:22)."""
from hoir_b200.sparse_optimizers import SparseLion  # noqa: F401
