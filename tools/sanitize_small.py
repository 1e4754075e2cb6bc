"""Dev tool: small invocations of every kernel family, meant to run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from bevfusion_b200.bev_pool import bev_pool, bev_pool_ext
from bevfusion_b200.spconv import ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
# drop-in bev_pool fwd+bwd with long intervals, odd channel count (generic kernel) and tuned width
for c in (80, 7):
    n, B, D, H, W = 5000, 2, 2, 9, 7
    coords = torch.from_numpy(np.stack([rng.integers(0, H, n), rng.integers(0, W, n), rng.integers(0, D, n),
                                        rng.integers(0, B, n)], 1)).to(dev)
    x = torch.randn(n, c, device=dev, requires_grad=True)
    out = bev_pool(x, coords, B, D, H, W)
    out.sum().backward()
# sparse conv fwd/bwd, all precisions, strided + subm
shape, Bn, n = [20, 18, 7], 2, 1500
vol = Bn * shape[0] * shape[1] * shape[2]
flat = rng.choice(vol, size=n, replace=False)
idx = np.stack([flat // (shape[0] * shape[1] * shape[2]), (flat // (shape[1] * shape[2])) % shape[0],
                (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
ti = torch.from_numpy(idx).to(dev)
for (cin, cout) in ((16, 32), (64, 64), (5, 16)):
    f = torch.randn(n, cin, device=dev)
    for subm, ks, st, pd in ((True, 3, 1, 1), (False, 3, 2, 1), (False, [1, 1, 3], [1, 1, 2], 0)):
        rb, _ = ops.get_rulebook(ti, Bn, shape, ks, st, pd, 1, 0, subm)
        kv = rb.nbr.shape[0]
        w = torch.randn(kv, cin, cout, device=dev)
        for prec in (0, 1, 2):
            o = ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=prec)
        ops.sparse_conv_backward(f, w, torch.randn_like(o), rb.nbr, precision=1)
        rb.pairs()
torch.cuda.synchronize()
print("sanitize_small ok")
