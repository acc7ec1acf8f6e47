"""EXECUTED drop-in test: the reference's own ``models.EAGCN`` (models.py:14-121), its own
``utils.weights_init`` and its own ``Dense`` / read-out / head code, with the five hot-path classes replaced by
``eagcn_b200.layers`` exactly as the one-line import swap at models.py:3 does -- forward + backward on the GPU --
against the unmodified reference classes run on the CPU with the same state_dict and inputs.

Needs a GPU and the reference modules (oracle/_ref on the GPU box, tools/make_oracle_ref.sh)."""
import pytest
import torch

from oracle import ref_loader
from tests.util import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

SWAPPED = ("GraphConv_Layer", "GraphConv_block", "GraphConv_base", "AFM_BatchNorm", "Ave_multi_view")


@pytest.mark.parametrize("molfp,training", [("sum", True), ("ave", False), ("pool", True)])
def test_reference_model_on_swapped_layers(molfp, training):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eagcn_b200 import layers as EL
    from eagcn_b200.data import make_batch
    dev = torch.device("cuda", 0)
    # --- the reference on the CPU (its own classes everywhere) ---
    with ref_loader.cpu_only():
        _, M_cpu, U_cpu = ref_loader.load()
        torch.manual_seed(0)
        ref = M_cpu.EAGCN(30, 24, *([8] * 5), *([12] * 5), 32, 16, 3, dropout=0.0, structure="Concate", molfp_mode=molfp)
        ref.apply(U_cpu.weights_init)                                      # train.py:302
        with torch.no_grad():
            for p in ref.parameters():                                     # AFM_BatchNorm.weight/.bias: uninitialised memory
                if not torch.isfinite(p).all():
                    p.fill_(0.0)
    # --- the reference's models.py with the import swap of INTEGRATION.md 1 (GPU: use_cuda is True) ---
    _, M_gpu, U_gpu = ref_loader.load()
    for name in SWAPPED:
        setattr(M_gpu, name, getattr(EL, name))                            # == "from eagcn_b200.layers import ..." at models.py:3
    model = M_gpu.EAGCN(30, 24, *([8] * 5), *([12] * 5), 32, 16, 3, dropout=0.0, structure="Concate", molfp_mode=molfp)
    assert type(model.layer1) is EL.GraphConv_Layer and type(model.layer1.block1.graph_conv) is EL.GraphConv_base
    model.cuda()                                                           # train.py:299-300
    model.apply(U_gpu.weights_init)                                        # class-name matching still finds the new classes
    model.load_state_dict(ref.state_dict(), strict=True)                   # identical state_dict keys / shapes
    ref.train(training); model.train(training)
    batch = make_batch(10, "tox21", seed=3, kb=30)
    dense = [torch.from_numpy(a) for a in batch.dense()]
    sizes = torch.from_numpy(batch.sizes)
    y_ref, atom_ref, g_ref = ref(*dense, sizes)
    y, atom, grep = model(*[t.to(dev) for t in dense], sizes.to(dev))       # models.py:96-121, unchanged code
    assert rel_err(atom, atom_ref) <= 4e-5                                  # 4 stacked layers at 1e-5 each
    assert rel_err(y.cpu(), y_ref) <= 1e-4 and rel_err(grep.cpu(), g_ref) <= 1e-4
    cot = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1))
    (y_ref * cot).sum().backward()
    (y * cot.to(dev)).sum().backward()
    named = dict(model.named_parameters())
    scale = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        got = named[k].grad
        assert got is not None, k
        denom = max(float(p.grad.abs().max()), 1e-3 * scale)
        if k.endswith("graph_conv.bias") and training:
            denom = scale                                                   # d bias through a train-mode BatchNorm: noise around 0
        assert float((got.cpu() - p.grad).abs().max()) / denom <= 2e-4, k
    # check_model.py:48-58 attribute paths
    assert model.layer4.block5.self_r.data.shape == (1,) and model.layer1.block1.att.weight.data.shape == (1, 30, 1, 1)
