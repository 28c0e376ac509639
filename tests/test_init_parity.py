"""the PRODUCT's ``initialize_network_polynomial_layers``
(/root/reference/high_order_implicit_representation/networks.py:57) -- bench.py builds the timed
network with it.  CPU: equal to the oracle's bit for bit under the same seeded generator, for every layer kind.
GPU: the initialised layers, evaluated by the CUDA kernels, ARE the straight lines  sum_i slope_i x_i + offset_i
(A.8 #4 through the kernels)."""
import pytest
import torch

import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import MLP_KW


@pytest.mark.parametrize("name", ["C1", "C2", "C3d", "C3f", "C4s", "poly", "C1n2"])
@pytest.mark.parametrize("slope,offset", [(1.0, 0.0), (0.7, 0.3)])
def test_product_init_equals_oracle_bit_for_bit(name, slope, offset):
    kw = MLP_KW[name]
    torch.manual_seed(11)
    ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    net.load_state_dict(ref.state_dict())                  # constructor init differs by RNG order: start equal
    before = {k: v.clone() for k, v in net.state_dict().items()}
    O.initialize_network_polynomial_layers(ref, max_slope=slope, max_offset=offset,
                                           generator=torch.Generator().manual_seed(5))
    H.initialize_network_polynomial_layers(net, max_slope=slope, max_offset=offset,
                                           generator=torch.Generator().manual_seed(5))
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    for (k, a), b in zip(net.state_dict().items(), ref.state_dict().values()):
        assert torch.equal(a, b), k
        if kw["layer_type"] == "fourier":
            assert torch.equal(a, before[k]), f"{k}: Fourier layers keep their constructor init"
        else:
            assert not torch.equal(a, before[k]), k
    for layer in net.high_order_layers():                  # node positions: the same fp32 table as the oracle's
        if kw["layer_type"] == "fourier":
            continue
        twin = O.high_order_fc_layers(kw["layer_type"], n=layer.n, in_features=1, out_features=1,
                                      segments=layer.segments)
        assert torch.equal(H.node_positions(layer), O.node_positions(twin))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,seg", [("continuous", 3, 100), ("continuous", 2, 50), ("discontinuous", 3, 7),
                                        ("polynomial", 4, 1), ("continuous", 5, 3)])
def test_initialised_cuda_layer_is_the_straight_line(kind, n, seg):
    torch.manual_seed(0)
    in_f, out_f = 6, 9
    net = H.HighOrderMLP(layer_type=kind, n=n, in_width=in_f, out_width=out_f, hidden_layers=0, hidden_width=12,
                         in_segments=seg, out_segments=seg, hidden_segments=seg)
    H.initialize_network_polynomial_layers(net, max_slope=1.0, max_offset=0.5,
                                           generator=torch.Generator().manual_seed(3))
    first = net.model[0]
    pos = H.node_positions(first)
    w = first.w.detach()
    slope = (w[:, :, -1] - w[:, :, 0]) / (pos[-1] - pos[0])
    offset = w[:, :, 0] - slope * pos[0]
    x = torch.rand(4096, in_f, generator=torch.Generator().manual_seed(4)) * 1.98 - 0.99
    want = x.double() @ slope.double().t() + offset.double().sum(1)
    got = first.cuda()(x.cuda())
    assert (got.double().cpu() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
