"""GPU diagnostic (not a test): per-tile error maps of the tcgen05 GEMM engine for a sweep of shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagcn_b200 import functional as EF

dev = torch.device("cuda", 0)


def nt(Mcap, T, N, K, BN_hint=None):
    g = torch.Generator().manual_seed(N * 7 + K)
    A = torch.randn(Mcap, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    m_dev = torch.tensor([T], dtype=torch.int32, device=dev)
    C = EF.gemm_nt(A, B, m_dev, 0)
    torch.cuda.synchronize()
    ref = A[:T].double() @ B.double().t()
    err = (C[:T].double() - ref).abs() / ref.abs().max()
    colerr = err.max(0).values
    rowerr = err.max(1).values
    bad_cols = (colerr > 1e-5).nonzero().flatten().tolist()
    bad_rows = (rowerr > 1e-5).nonzero().flatten().tolist()
    def rng(v):
        if not v: return "-"
        out, s, p = [], v[0], v[0]
        for x in v[1:]:
            if x != p + 1: out.append((s, p)); s = x
            p = x
        out.append((s, p))
        return ",".join(f"{a}-{b}" for a, b in out[:12]) + ("..." if len(out) > 12 else "")
    print(f"NT M={T}/{Mcap} N={N} K={K}: max {float(err.max()):.2e}  bad cols [{rng(bad_cols)}] bad rows [{rng(bad_rows)}]", flush=True)


def tn(Kcap, T, M, N):
    g = torch.Generator().manual_seed(M * 5 + N)
    A = torch.randn(Kcap, M, generator=g); B = torch.randn(Kcap, N, generator=g)
    A[T:] = 0; B[T:] = 0
    A, B = A.to(dev), B.to(dev)
    k_dev = torch.tensor([T], dtype=torch.int32, device=dev)
    C = EF.gemm_tn(A, B, k_dev, 0)
    torch.cuda.synchronize()
    ref = A.double().t() @ B.double()
    err = (C.double() - ref).abs() / ref.abs().max()
    nz = float((C != 0).float().mean())
    print(f"TN K={T}/{Kcap} M={M} N={N}: max {float(err.max()):.2e} mean {float(err.mean()):.2e} nonzero-frac {nz:.2f} "
          f"C[0,:4]={C[0,:4].tolist()} ref={ref[0,:4].tolist()}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "nt"):
        for (Mcap, T, N, K) in [(128, 128, 80, 32), (128, 128, 80, 64), (128, 128, 80, 96), (128, 128, 80, 700),
                                (896, 860, 400, 700), (128, 128, 400, 64), (128, 128, 96, 128), (128, 128, 112, 128),
                                (128, 128, 144, 128), (128, 128, 64, 128), (128, 128, 176, 128), (128, 128, 240, 128),
                                (128, 128, 48, 128), (256, 256, 160, 256)]:
            try:
                nt(Mcap, T, N, K)
            except Exception as e:
                print("NT", (Mcap, T, N, K), "EXC", e, flush=True)
    if which in ("all", "tn"):
        for (Kcap, T, M, N) in [(128, 128, 32, 32), (128, 128, 128, 32), (128, 128, 32, 128), (256, 200, 24, 400),
                                (512, 511, 400, 700), (4864, 4853, 400, 700)]:
            try:
                tn(Kcap, T, M, N)
            except Exception as e:
                print("TN", (Kcap, T, M, N), "EXC", e, flush=True)
