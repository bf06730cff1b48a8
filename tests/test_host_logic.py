"""CPU tests of the host-side mirror: constructor / state_dict contract, EMA pre-processing,
factory behaviour, scatter helpers -- everything that does not need a kernel."""
import types

import pytest
import torch

from se3conv3d_b200.layers import PNEConvLayerRotEquiv, PNEConvLayerRotEquivFactory, PreProcessModule
from se3conv3d_b200.pc import BQNeighborhood, all_index_combinations
from se3conv3d_b200 import scatter


def test_state_dict_contract_and_init_ranges():
    torch.manual_seed(2)
    layer = PNEConvLayerRotEquiv(9, 32, 64, 32, "mlp_gelu")
    sd = layer.state_dict()
    assert list(sd.keys()) == ["proj_axes_", "proj_biases_", "conv_weights_", "norm_neigh_dist_", "norm_num_neighs_"]
    assert tuple(sd["proj_axes_"].shape) == (9, 32) and tuple(sd["conv_weights_"].shape) == (32, 32, 64)
    assert float(sd["proj_axes_"].abs().max()) <= (1 / 9) ** 0.5
    assert float(sd["conv_weights_"].abs().max()) <= (1 / (32 * 32)) ** 0.5
    assert float(sd["proj_biases_"].abs().max()) == 0 and float(sd["norm_num_neighs_"]) == 0
    # same seed, same construction order as the reference -> same tensors as torch's own draws
    torch.manual_seed(2)
    a = torch.empty(9, 32).uniform_(-(1 / 9) ** 0.5, (1 / 9) ** 0.5)
    w = torch.empty(32, 32, 64).uniform_(-(1 / 1024) ** 0.5, (1 / 1024) ** 0.5)
    assert torch.equal(a, sd["proj_axes_"]) and torch.equal(w, sd["conv_weights_"])


def test_factory_and_cache_api():
    f = PNEConvLayerRotEquivFactory(p_dims=9, p_num_basis=32, p_pne_type="mlp_gelu")
    l1, l2 = f.create_conv_layer(1, 32), f.create_conv_layer(32, 64)
    assert f.conv_list_ == [l1, l2] and l2.feat_output_size_ == 64
    f.update_parameters(num_basis=16)
    assert f.num_basis_ == 16
    PNEConvLayerRotEquiv.rot_tensor_cache["x"] = 1
    PNEConvLayerRotEquiv.empty_rot_tenors_cache()
    assert PNEConvLayerRotEquiv.rot_tensor_cache == {}
    assert PNEConvLayerRotEquiv.rel_rot_type == "6D"


def test_preprocess_ema_updates_buffers():
    class Net(PreProcessModule):
        def __init__(self):
            super().__init__()
            self.convs = torch.nn.ModuleList([PNEConvLayerRotEquiv(9, 4, 4, 32, "mlp_gelu")])

    net = Net()
    conv = net.convs[0]
    conv.__compute_convolution__ = types.MethodType(lambda self, *a: None, conv)
    neigh = BQNeighborhood.__new__(BQNeighborhood)
    neigh.radius_ = 0.5
    neigh.neighbors_ = torch.zeros(40, 2, dtype=torch.int64)
    neigh.start_ids_ = torch.zeros(10, dtype=torch.int32)
    conv(None, None, None, neigh)
    assert float(conv.norm_neigh_dist_) == 0.0           # not pre-processing: untouched
    net.start_pre_process()
    assert conv.pre_process_
    conv(None, None, None, neigh)
    assert abs(float(conv.norm_neigh_dist_) - 0.1 * 2.0) < 1e-7
    assert abs(float(conv.norm_num_neighs_) - 0.1 * 0.25) < 1e-7
    conv(None, None, None, neigh)
    assert abs(float(conv.norm_neigh_dist_) - (0.9 * 0.2 + 0.2)) < 1e-6
    net.end_pre_process()
    assert not conv.pre_process_


def test_scatter_helpers():
    src = torch.tensor([[1.0, 2.0], [3.0, 5.0], [-1.0, 0.0], [7.0, 7.0]])
    idx = torch.tensor([0, 2, 0, 2])
    assert torch.equal(scatter.scatter_add(src, idx), torch.tensor([[0.0, 2.0], [0.0, 0.0], [10.0, 12.0]]))
    assert torch.equal(scatter.scatter_mean(src, idx), torch.tensor([[0.0, 1.0], [0.0, 0.0], [5.0, 6.0]]))
    assert torch.equal(scatter.scatter_max(src, idx)[0], torch.tensor([[1.0, 2.0], [0.0, 0.0], [7.0, 7.0]]))
    assert torch.equal(scatter.scatter_min(src, idx)[0], torch.tensor([[-1.0, 0.0], [0.0, 0.0], [3.0, 5.0]]))
    assert all_index_combinations(2, 3).tolist() == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2]]


def test_block_modules_state_dict_and_drop_path():
    """Row f2: ResNetFormer / BatchNormPC / SkipConnection / DropPathPC keep the reference's parameter names
    (layers/ResNetFormer.py:37-52, SkipConnection.py:27-29, BatchNormPC.py:21) and drop-path semantics."""
    import types
    from se3conv3d_b200.layers import (ResNetFormer, BatchNormPC, SkipConnection, DropPathPC,
                                       PNEConvLayerRotEquivFactory)
    fact = PNEConvLayerRotEquivFactory(9, 32, "mlp_gelu")
    blk = ResNetFormer(16, 24, fact, BatchNormPC, 0.1)
    keys = set(blk.state_dict().keys())
    for k in ("spatial_conv_.proj_axes_", "spatial_conv_.conv_weights_", "norm_1_.layer_.weight",
              "norm_2_.layer_.running_mean", "linear_1_.weight", "linear_2_.bias", "skip_path_1_.gamma_",
              "skip_path_2_.gamma_", "skip_conv_.weight"):
        assert k in keys, k
    assert blk.linear_1_.weight.shape == (32, 16) and blk.linear_2_.weight.shape == (24, 32)
    assert float(blk.skip_path_1_.gamma_[0, 0]) == pytest.approx(1e-6)
    assert blk.norm_1_.layer_.momentum == 0.2 and len(fact.conv_list_) == 1
    # drop path: identity in eval mode; in training whole batch items are zeroed and the rest rescaled
    pc = types.SimpleNamespace(batch_size_=torch.tensor(4), batch_ids_=torch.tensor([0, 0, 1, 2, 3, 3], dtype=torch.int32))
    dp = DropPathPC(0.5)
    x = torch.ones(6, 3)
    dp.eval()
    assert torch.equal(dp(x, pc), x)
    dp.train()
    torch.manual_seed(0)
    y = dp(x, pc)
    assert set(y.unique().tolist()) <= {0.0, 2.0}
    assert torch.equal(y[0], y[1]) and torch.equal(y[4], y[5])
    sk = SkipConnection(0.0, 3)
    assert torch.allclose(sk(x, 2 * x, pc), 2 * x + 1e-6)


def test_lazy_attrs_materialise_once_and_keep_plain_attribute_semantics():
    """LazyAttrs (the windows of a fused hierarchy arena): a thunk runs on first access only, the value then is a
    plain attribute; names without a thunk raise AttributeError (so hasattr / getattr-with-default keep working)."""
    from se3conv3d_b200._lib import LazyAttrs

    class Obj(LazyAttrs):
        pass

    calls = []
    o = Obj()
    o.plain = 1
    o._lazy = {"window": lambda: calls.append(1) or "tensor"}
    assert o.plain == 1 and not hasattr(o, "missing") and getattr(o, "missing", None) is None
    assert o.window == "tensor" and o.window == "tensor" and calls == [1]
    assert "window" in o.__dict__ and "window" not in o._lazy
    o.window = "replaced"
    assert o.window == "replaced"
    bare = Obj()                      # objects built by the per-object path carry no thunks at all
    with pytest.raises(AttributeError):
        bare.window
