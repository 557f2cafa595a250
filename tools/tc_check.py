"""First-contact check of the tcgen05 layer kernel (run under `timeout`): prints errors instead of asserting."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_tc_gpu import DEV, tc_linear, tf32_trunc  # noqa: E402

for (M, K, act, acc) in [(128, 128, False, False), (128, 32, False, False), (256, 128, False, False), (1024, 128, True, False), (5000, 72, False, False),
                         (128 * 300, 128, True, True)]:
    g = torch.Generator().manual_seed(M + K)
    X = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(128, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(128, generator=g).to(DEV)
    a = (1 + 0.2 * torch.randn(K, generator=g)).to(DEV) if act else None
    b = (0.3 * torch.randn(K, generator=g)).to(DEV) if act else None
    Y0 = torch.randn(M, 128, generator=g).to(DEV) if acc else None
    Y, s = tc_linear(X, W, bias, a, b, Y0)
    Xa = torch.nn.functional.softplus(X * a + b) if act else X
    ref = Xa.double() @ W.double().t() + bias.double() + (Y0.double() if acc else 0)
    ref_t = tf32_trunc(Xa).double() @ tf32_trunc(W).double().t() + bias.double() + (Y0.double() if acc else 0)
    e = (Y.double() - ref).abs()
    print(f"M={M} K={K} act={act} acc={acc}: max|err| vs fp64 {e.max().item():.3e} (ref max {ref.abs().max().item():.2f}); vs tf32-trunc "
          f"{(Y.double() - ref_t).abs().max().item():.3e}; stats err {(s[0] - Y.double().sum(0)).abs().max().item():.2e}", flush=True)
    if e.max().item() > 0.05:
        bad = (e > 0.05).nonzero()
        print("   bad rows (first 10):", bad[:10].tolist(), " #bad", bad.shape[0], " rows with any bad:", bad[:, 0].unique().numel(),
              " cols with any bad:", bad[:, 1].unique().numel(), flush=True)
        print("   Y[0,:8]", Y[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist(), flush=True)
