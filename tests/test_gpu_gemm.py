"""GPU: the projection GEMM engines in isolation against a float64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.mark.parametrize("engine", [0, 1, 16, 32])
@pytest.mark.parametrize("Mcap,T,N,K", [
    (128, 128, 64, 32), (256, 200, 400, 24), (512, 511, 700, 400), (384, 300, 24, 400), (256, 129, 176, 40),
    (1024, 1000, 1400, 700), (128, 1, 16, 8), (4864, 4853, 400, 700), (4864, 4853, 700, 400),
])
def test_gemm_nt_vs_float64(engine, Mcap, T, N, K):
    """engine 0: tcgen05 with the k-block the pipeline model picks; 16 / 32: tcgen05 forced to 64-byte (SWIZZLE_64B) /
    128-byte (SWIZZLE_128B) k-blocks; 1: FFMA."""
    from eagcn_b200 import functional as EF, _lib
    dev = _cuda()
    if engine in (16, 32):
        _lib.lib().eagcn_set_tc_bk(engine)
        try:
            _gemm_nt_case(EF, dev, 0, Mcap, T, N, K)
        finally:
            _lib.lib().eagcn_set_tc_bk(0)
        return
    _gemm_nt_case(EF, dev, engine, Mcap, T, N, K)


def _gemm_nt_case(EF, dev, engine, Mcap, T, N, K):
    g = torch.Generator(device="cpu").manual_seed(Mcap + N + K)
    A = torch.randn(Mcap, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    m_dev = torch.tensor([T], dtype=torch.int32, device=dev)
    C = EF.gemm_nt(A, B, m_dev, engine)
    torch.cuda.synchronize()
    ref = (A[:T].double() @ B.double().t())
    err = float((C[:T].double() - ref).abs().max()) / float(ref.abs().max())
    print(f"engine {engine} M={T} N={N} K={K}: err {err:.2e}")
    # fp32-class accuracy: FFMA ~1e-7 of max; 3xTF32 tensor-core a few 1e-6 (truncating fp32 accumulation in
    # the tensor core, K/8 steps) -- inside the 1e-5 layer-level parity bar
    assert err <= (6e-6 if engine == 0 else 2e-6), (engine, err)
    assert float(C[T:].abs().max()) == 0.0 if T < Mcap else True


def test_gemm_unaligned_layout_is_rejected_not_wrong():
    from eagcn_b200 import functional as EF
    from eagcn_b200._lib import EagcnError
    dev = _cuda()
    A = torch.randn(128, 9, device=dev)
    B = torch.randn(20, 9, device=dev)
    m_dev = torch.tensor([128], dtype=torch.int32, device=dev)
    with pytest.raises(EagcnError):
        EF.gemm_nt(A, B, m_dev, engine=0)          # 36-byte rows: TMA cannot address them
    C = EF.gemm_nt(A, B, m_dev, engine=1)
    ref = A.double() @ B.double().t()
    assert float((C.double() - ref).abs().max()) / float(ref.abs().max()) <= 1e-6


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("Kcap,T,M,N", [
    (128, 128, 32, 32), (256, 200, 24, 400), (512, 511, 400, 700), (4864, 4853, 400, 700), (1024, 1000, 700, 1400),
    (256, 33, 128, 64), (384, 300, 100, 36),
])
def test_gemm_tn_vs_float64(engine, Kcap, T, M, N):
    """dW = H^T Q: MN-major tcgen05 operands + split-K (engine 0), FFMA (engine 1)."""
    from eagcn_b200 import functional as EF
    dev = _cuda()
    g = torch.Generator(device="cpu").manual_seed(Kcap + M + N)
    A = torch.randn(Kcap, M, generator=g)
    B = torch.randn(Kcap, N, generator=g)
    A[T:] = 0; B[T:] = 0                                   # slack rows are zero by contract
    A, B = A.to(dev), B.to(dev)
    k_dev = torch.tensor([T], dtype=torch.int32, device=dev)
    C = EF.gemm_tn(A, B, k_dev, engine)
    torch.cuda.synchronize()
    ref = A.double().t() @ B.double()
    err = float((C.double() - ref).abs().max()) / float(ref.abs().max())
    print(f"TN engine {engine} K={T} M={M} N={N}: err {err:.2e}")
    assert err <= (6e-6 if engine == 0 else 2e-6), (engine, err)


def test_dense_mm_autograd():
    from eagcn_b200 import functional as EF
    dev = _cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(256, 300, generator=g).to(dev).requires_grad_(True)
    W = torch.randn(300, 64, generator=g).to(dev).requires_grad_(True)
    R = torch.randn(256, 64, generator=g).to(dev)
    (EF.dense_mm(x, W) * R).sum().backward()
    xr = x.detach().double().requires_grad_(True); Wr = W.detach().double().requires_grad_(True)
    ((xr @ Wr) * R.double()).sum().backward()
    assert float((x.grad.double() - xr.grad).abs().max() / xr.grad.abs().max()) <= 2e-6
    assert float((W.grad.double() - Wr.grad).abs().max() / Wr.grad.abs().max()) <= 2e-6


@pytest.mark.parametrize("M,N,K,tA,tB", [
    (256, 256, 700, False, False),      # den1 forward (Tox21 head)
    (256, 64, 256, False, False),       # den2 forward
    (256, 12, 64, False, False),        # den3 forward
    (256, 700, 256, False, True),       # dX of den1
    (700, 256, 256, True, False),       # dW of den1
    (64, 12, 256, True, False),         # dW of den3
    (256, 1, 64, False, False),         # regression head: one output column (scalar edge path)
    (1, 64, 256, True, False),
    (37, 45, 131, False, False),        # nothing aligned
    (37, 45, 131, True, True),
    (33, 31, 7, False, True),           # K shorter than one vector group
    (512, 1400, 256, False, True),      # wide read-out (4-layer model)
    (256, 256, 1400, False, False),     # den1 forward of the 4-layer model: 11 split-K slabs
    (40, 33, 1, False, False),          # K = 1: both strides of an operand are 1
    (40, 33, 1, True, True),
    (64, 64, 300, True, True),          # three slabs through the two-stage pipeline, both operands k-contiguous
])
def test_mm_tile_vs_float64(M, N, K, tA, tB):
    """eagcn_mm_tile (mm_tile.cu, the dense layers of the head: layers.py:382-388) against a float64 product; the
    in-kernel split-K combine is bit-reproducible and leaves its ticket array zero."""
    from eagcn_b200 import functional as EF
    dev = _cuda()
    g = torch.Generator().manual_seed(M * 131 + N * 7 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(dev)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(dev)
    C, _ = EF._mm_tile(A, tA, B, tB)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err
    for _ in range(3):
        C2, _ = EF._mm_tile(A, tA, B, tB)
        assert torch.equal(C, C2)
    for t in EF._tickets.values():
        assert int(t.abs().sum()) == 0


def test_mm_tile_strided_operands():
    """lda / ldb larger than the logical width (views into wider buffers)."""
    from eagcn_b200 import functional as EF, _lib
    from eagcn_b200._lib import check, lib, ptr
    dev = _cuda()
    g = torch.Generator().manual_seed(5)
    Abig = torch.randn(100, 90, generator=g).to(dev)
    Bbig = torch.randn(70, 50, generator=g).to(dev)
    M, K, N = 100, 70, 44
    C = torch.empty(M, N, device=dev)
    L = lib()
    nbytes = int(L.eagcn_mm_tile_workspace_bytes(M, N, K))
    ws = torch.empty(max(nbytes // 4, 1), device=dev)
    tk = torch.zeros(int(L.eagcn_mm_tile_tickets(M, N)), dtype=torch.int32, device=dev)
    import ctypes
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(L.eagcn_mm_tile(ptr(Abig), 90, 0, ptr(Bbig), 50, 0, ptr(C), M, N, K, ptr(ws), nbytes, ptr(tk), st), "mm_tile")
    ref = Abig[:, :K].double() @ Bbig[:, :N].double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) <= 2e-6
    # invalid arguments come back as status codes, never as a crash
    assert L.eagcn_mm_tile(None, 90, 0, ptr(Bbig), 50, 0, ptr(C), M, N, K, ptr(ws), nbytes, ptr(tk), st) == _lib.lib().eagcn_mm_tile(None, 90, 0, ptr(Bbig), 50, 0, ptr(C), M, N, K, ptr(ws), nbytes, ptr(tk), st) < 0
    assert L.eagcn_mm_tile(ptr(Abig), 10, 0, ptr(Bbig), 50, 0, ptr(C), M, N, K, ptr(ws), nbytes, ptr(tk), st) < 0


@pytest.mark.parametrize("overlap", [False, True])
def test_dense_mm_tile_autograd(overlap):
    from eagcn_b200 import functional as EF
    dev = _cuda()
    old = EF.Overlap.enabled
    EF.Overlap.enabled = overlap
    try:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(256, 300, generator=g).to(dev).requires_grad_(True)
        W = torch.randn(300, 64, generator=g).to(dev).requires_grad_(True)
        R = torch.randn(256, 64, generator=g).to(dev)
        (EF.dense_mm(x, W, "tile") * R).sum().backward()
        torch.cuda.synchronize()
        xr = x.detach().double().requires_grad_(True); Wr = W.detach().double().requires_grad_(True)
        ((xr @ Wr) * R.double()).sum().backward()
        assert float((x.grad.double() - xr.grad).abs().max() / xr.grad.abs().max()) <= 2e-6
        assert float((W.grad.double() - Wr.grad).abs().max() / Wr.grad.abs().max()) <= 2e-6
        # accumulation into an existing .grad (the join must then precede the accumulate kernel)
        g0 = W.grad.clone()
        (EF.dense_mm(x, W, "tile") * R).sum().backward()
        torch.cuda.synchronize()
        assert float((W.grad - 2 * g0).abs().max() / g0.abs().max()) <= 1e-6
    finally:
        EF.Overlap.enabled = old


def test_tf32_single_pass_mode_is_opt_in_and_coarser():
    """eagcn_set_tc_passes(1): ONE TF32 tensor-core pass on the raw operands (BASELINE.json's reduced-precision
    configuration, reported beside the strict mode).  It must really be a different engine (error ~1e-4..1e-3 against
    float64, where the default 3xTF32 stays <= 6e-6), must leave the default untouched, and a whole layer must still
    agree with the strict mode to TF32 accuracy (forward and gradients)."""
    from eagcn_b200 import functional as EF, layers as EL, _lib
    from eagcn_b200.data import make_batch
    dev = _cuda()
    L = _lib.lib()
    assert L.eagcn_get_tc_passes() == 3
    g = torch.Generator().manual_seed(9)
    A = torch.randn(640, 400, generator=g).to(dev)
    B = torch.randn(700, 400, generator=g).to(dev)
    m = torch.tensor([600], dtype=torch.int32, device=dev)
    ref = (A.double() @ B.double().t())
    ref[600:] = 0
    errs = {}
    batch = make_batch(48, "tox21", seed=2, n_afeat=64)
    torch.manual_seed(1)
    layer = EL.GraphConv_Layer(64, 30, 40, 40, 24, 16, 8, dropout=0.0, structure="Concate").to(dev).train()
    ins = [torch.from_numpy(a).to(dev) for a in batch.dense()]
    outs = {}
    try:
        for name in ("fp32x3", "tf32"):
            EF.set_tc_precision(name)
            C = EF.gemm_nt(A, B, m, engine=0)
            errs[name] = float((C.double() - ref).abs().max() / ref.abs().max())
            for p in layer.parameters():
                p.grad = None
            afm = ins[1].clone().requires_grad_(True)
            x, _ = layer(ins[0], afm, *ins[2:])
            x.square().sum().backward()
            outs[name] = (x.detach().clone(), afm.grad.clone(), layer.block1.graph_conv.weight.grad.clone())
    finally:
        EF.set_tc_precision("fp32x3")
    assert errs["fp32x3"] <= 6e-6, errs
    assert 2e-5 < errs["tf32"] <= 5e-3, errs
    for a, b in zip(outs["tf32"], outs["fp32x3"]):
        e = float((a - b).abs().max() / b.abs().max())
        assert 1e-7 < e <= 2e-2, e
    assert L.eagcn_get_tc_passes() == 3


@pytest.mark.parametrize("Mcap,T,N,K", [
    (128, 128, 64, 32), (512, 511, 700, 400), (256, 129, 176, 40), (4864, 4853, 400, 700), (4864, 4853, 700, 400),
    (384, 300, 24, 400), (128, 1, 16, 8),
])
def test_gemm_activation_operand_in_tensor_memory(Mcap, T, N, K):
    """eagcn_set_tc_a_tmem: the hi / lo split of the activation operand written to TENSOR MEMORY (tcgen05.st, MMA with A
    from TMEM) against the shared-memory hi / lo copies.  Same split values, same MMA order per accumulator: the K-major
    product is BIT-IDENTICAL; the split-K product differs only through the number of K splits the tile shape implies."""
    from eagcn_b200 import functional as EF, _lib
    dev = _cuda()
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(Mcap + N + K + 1)
    A = torch.randn(Mcap, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    m_dev = torch.tensor([T], dtype=torch.int32, device=dev)
    # TN operands: [Kcap = Mcap rows, M = K columns] and [Kcap, N]
    At = torch.randn(Mcap, K, generator=g); Bt = torch.randn(Mcap, N, generator=g)
    At[T:] = 0; Bt[T:] = 0
    At, Bt = At.to(dev), Bt.to(dev)
    assert L.eagcn_get_tc_a_tmem() == 3                      # the default
    out = {}
    try:
        for mask in (0, 3):
            assert L.eagcn_set_tc_a_tmem(mask) == 0
            out[mask] = (EF.gemm_nt(A, B, m_dev, 0).clone(),
                         EF.gemm_tn(At, Bt, m_dev, 0).clone() if min(K, N) >= 32 else None)
            torch.cuda.synchronize()
    finally:
        L.eagcn_set_tc_a_tmem(3)
    assert torch.equal(out[0][0], out[3][0])
    ref = At.double().t() @ Bt.double()
    for mask in (0, 3):
        if out[mask][1] is not None:
            assert float((out[mask][1].double() - ref).abs().max()) / float(ref.abs().max()) <= 6e-6, mask
    assert L.eagcn_set_tc_a_tmem(4) != 0
