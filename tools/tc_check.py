"""Quick GPU check of the tcgen05 sparse-conv path against the CPU oracle (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import oracle
from bevfusion_b200.spconv import ops

dev = torch.device("cuda:0")
def random_sparse(n, shape, B, seed):
    rng = np.random.default_rng(seed)
    vol = B * shape[0] * shape[1] * shape[2]
    flat = rng.choice(vol, size=n, replace=False)
    z = flat % shape[2]; y = (flat // shape[2]) % shape[1]
    x = (flat // (shape[2] * shape[1])) % shape[0]; b = flat // (shape[2] * shape[1] * shape[0])
    return np.stack([b, x, y, z], 1).astype(np.int32)

shape, B, n = [40, 36, 11], 2, 6000
for (cin, cout) in [(32, 32), (16, 16), (64, 64), (128, 128), (16, 32), (64, 128)]:
    idx = random_sparse(n, shape, B, seed=cin + cout)
    rng = np.random.default_rng(7)
    feat = rng.standard_normal((n, cin)).astype(np.float32)
    W = (rng.standard_normal((3, 3, 3, cin, cout)) / np.sqrt(cin * 9)).astype(np.float32)
    gold, gids, _ = oracle.sparse_conv(feat, idx, B, shape, W, [3]*3, [1]*3, [1]*3, [1]*3, True, acc64=True)
    rb, _ = ops.get_rulebook(torch.from_numpy(idx).to(dev), B, shape, 3, 1, 1, 1, 0, True)
    f, w = torch.from_numpy(feat).to(dev), torch.from_numpy(W).to(dev)
    for prec in (0, 1, 2, 3):
        out = ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=prec)
        torch.cuda.synchronize()
        err = np.abs(out.cpu().numpy() - gold).max() / np.abs(gold).max()
        print(f"cin {cin:4d} cout {cout:4d} prec {prec}: rel err {err:.3e}", flush=True)
