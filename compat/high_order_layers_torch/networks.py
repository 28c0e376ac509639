"""``from high_order_layers_torch.networks import *`` (networks.py:8): HighOrderMLP,
initialize_network_polynomial_layers."""
from hoir_b200.networks import HighOrderMLP, initialize_network_polynomial_layers  # noqa: F401
