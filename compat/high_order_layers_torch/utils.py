"""``from high_order_layers_torch.utils import positions_from_mesh`` (single_image_dataset.py:13)."""
from hoir_b200.utils import positions_from_mesh  # noqa: F401
