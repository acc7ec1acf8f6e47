"""CPU: vectorised device-side losses vs the loop-for-loop restatement of utils.weight_tensor / train.py:326-331."""
import pytest
import torch

from eagcn_b200 import losses
from oracle import eagcn_oracle as O


def test_weighted_bce_matches_reference_loop():
    g = torch.Generator().manual_seed(0)
    B, T = 37, 12
    labels = torch.randint(0, 2, (B, T), generator=g).float()
    missing = torch.rand(B, T, generator=g) < 0.3
    labels[missing] = -1.0
    labels[0, 0] = float("nan")
    outputs = torch.randn(B, T, generator=g, requires_grad=True)
    weights = {j: [5000.0 / (10 + 3 * j), 5000.0 / (200 + j)] for j in range(T)}
    table = losses.bce_weight_table(weights, T)
    w_vec = losses.label_weights(table, labels).reshape(-1)
    w_ref = O.weight_tensor_loop(weights, labels)
    assert torch.equal(w_vec, w_ref)                            # weights: exact
    loss = losses.weighted_bce_with_logits(outputs, labels, table)
    ref = O.bce_loss_loop(outputs.detach().clone().requires_grad_(True), labels, weights)
    assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
    loss.backward()
    assert torch.isfinite(outputs.grad).all()
    assert float(outputs.grad[missing].abs().max()) == 0.0       # missing labels carry no gradient


def test_mse():
    y, t = torch.randn(9, 1), torch.randn(9)
    assert torch.allclose(losses.mse(y, t), ((y.view(-1) - t) ** 2).mean())


@pytest.mark.gpu
def test_losses_on_device_vs_reference_loop():
    """The device-side losses (no host loop, no synchronisation) against the loop-for-loop restatement of
    utils.weight_tensor / train.py:321-331 (itself pinned to the live reference, tests/test_reference_pins.py)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    B, T = 256, 12                                              # Tox21 batch of the benchmark: 3 072 (molecule, task) pairs
    labels = torch.randint(0, 2, (B, T), generator=g).float()
    missing = torch.rand(B, T, generator=g) < 0.25
    labels[missing] = -1.0
    labels[3, 7] = float("nan")
    missing[3, 7] = True
    outputs = torch.randn(B, T, generator=g)
    weights = {j: [5000.0 / (10 + 3 * j), 5000.0 / (200 + j)] for j in range(T)}
    table = losses.bce_weight_table(weights, T, device=dev)
    assert torch.equal(losses.label_weights(table, labels.to(dev)).reshape(-1).cpu(), O.weight_tensor_loop(weights, labels))
    o_dev = outputs.to(dev).requires_grad_(True)
    loss = losses.weighted_bce_with_logits(o_dev, labels.to(dev), table)
    loss.backward()
    o_ref = outputs.clone().requires_grad_(True)
    ref = O.bce_loss_loop(o_ref, labels, weights)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((o_dev.grad.cpu() - o_ref.grad).abs().max()) <= 1e-5 * float(o_ref.grad.abs().max())
    assert float(o_dev.grad[missing.to(dev)].abs().max()) == 0.0
    # regression branch (train.py:321-325)
    y, t = torch.randn(B, 1, generator=g), torch.randn(B, generator=g)
    assert abs(float(losses.mse(y.to(dev), t.to(dev))) - float(((y.view(-1) - t) ** 2).mean())) <= 1e-6
