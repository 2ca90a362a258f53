"""GPU diagnostic: tcgen05 vs mma.sync window attention against the fp32 SIMT path on the same data
(relative Frobenius error per output; localises errors per 128-row tile of the window)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valor_b200 import kernels as K


def run(grid, win, shift, heads, dtype, backend, qkv32, do32, table):
    hd = 32
    geom = (grid, win, shift, (8, 7, 7), heads, hd, hd ** -0.5)
    qkv, do = qkv32.to(dtype), do32.to(dtype)
    o, lse = K.window_attn_fwd(qkv, table, *geom, backend=backend)
    dt = torch.zeros_like(table)
    dqkv = K.window_attn_bwd(qkv, o, do, lse, table, dt, *geom, backend=backend)
    torch.cuda.synchronize()
    return o.float(), lse, dqkv.float(), dt


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


for grid, win, shift, heads in [((2, 8, 14, 14), (8, 7, 7), (0, 3, 3), 4), ((2, 8, 14, 14), (8, 7, 7), (0, 0, 0), 4),
                                ((2, 4, 14, 14), (4, 7, 7), (0, 3, 3), 4), ((2, 2, 14, 14), (2, 7, 7), (0, 3, 3), 4)]:
    C = heads * 32
    tokens = grid[0] * grid[1] * grid[2] * grid[3]
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv32 = torch.randn(tokens, 3 * C, device="cuda", generator=g).bfloat16().float()
    do32 = torch.randn(tokens, C, device="cuda", generator=g).bfloat16().float()
    table = torch.randn(15 * 13 * 13, heads, device="cuda", generator=g) * 0.5
    ref = run(grid, win, shift, heads, torch.float32, K.BACKEND_SIMT, qkv32, do32, table)
    out = {"grid": grid, "win": win, "shift": shift}
    for name, be in (("tcgen05", K.BACKEND_TENSOR), ("mma_sync", K.BACKEND_MMA_SYNC)):
        got = run(grid, win, shift, heads, torch.bfloat16, be, qkv32, do32, table)
        out[name] = {"o": rel(got[0], ref[0]), "lse": rel(got[1], ref[1]), "dq": rel(got[2][:, :C], ref[2][:, :C]),
                     "dk": rel(got[2][:, C:2 * C], ref[2][:, C:2 * C]), "dv": rel(got[2][:, 2 * C:], ref[2][:, 2 * C:]),
                     "dtable": rel(got[3], ref[3])}
        if name == "tcgen05":   # worst rows
            err = (got[2] - ref[2]).abs().max(dim=1).values
            out["tcgen05_worst_rows"] = err.topk(5).indices.tolist()
            out["tcgen05_worst_vals"] = [round(v, 4) for v in err.topk(5).values.tolist()]
    print(json.dumps(out), flush=True)
