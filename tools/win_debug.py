"""Localise window-attention backward mismatches: per output part (dQ/dK/dV), per window and per local token."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cpu_backend as R  # noqa: E402
from valor_b200 import kernels as K  # noqa: E402


def rnd(*shape, seed=0, scale=1.0, bf=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * scale
    return x.to(torch.bfloat16).float() if bf else x


def main():
    grid, win, shift, heads = (1, 8, 14, 7), (8, 7, 7), (0, 3, 0), 2
    if len(sys.argv) > 1:
        grid, win, shift, heads = eval(sys.argv[1])
    hd, C, cfg = 32, heads * 32, (8, 7, 7)
    tokens = grid[0] * grid[1] * grid[2] * grid[3]
    qkv, table, do = rnd(tokens, 3 * C, seed=1), rnd(15 * 13 * 13, heads, seed=2, scale=0.5, bf=False), rnd(tokens, C, seed=3)
    sc = hd ** -0.5
    o_r, lse_r = R.window_attn_fwd(qkv, table, grid, win, shift, cfg, heads, hd, sc)
    dt_r = torch.zeros_like(table)
    dqkv_r = R.window_attn_bwd(qkv, o_r, do, lse_r, table, dt_r, grid, win, shift, cfg, heads, hd, sc)
    q = qkv.cuda().bfloat16()
    o, lse = K.window_attn_fwd(q, table.cuda(), grid, win, shift, cfg, heads, hd, sc)
    dt = torch.zeros_like(table).cuda()
    dqkv = K.window_attn_bwd(q, o, do.cuda().bfloat16(), lse, table.cuda(), dt, grid, win, shift, cfg, heads, hd, sc)
    err = (dqkv.float().cpu() - dqkv_r).abs()
    print("o err", (o.float().cpu() - o_r).abs().max().item(), "lse err", (lse.cpu() - lse_r).abs().max().item())
    print("dtable err", (dt.cpu() - dt_r).abs().max().item(), "scale", dt_r.abs().max().item())
    B, D, H, W = grid
    for part, name in enumerate(["dQ", "dK", "dV"]):
        for h in range(heads):
            e = err[:, part * C + h * hd: part * C + (h + 1) * hd].amax(dim=1).view(B, D, H, W)
            print(name, "head", h, "max", e.max().item(), "scale", dqkv_r[:, part * C + h * hd: part * C + (h + 1) * hd].abs().max().item())
            bad = (e > 0.05).nonzero()
            print("   bad tokens:", bad.shape[0], bad[:12].tolist())
            # local query index inside the (shifted) window
            loc = []
            for b_, d_, h_, w_ in bad.tolist():
                cd, chh, cw = (d_ - shift[0]) % D, (h_ - shift[1]) % H, (w_ - shift[2]) % W
                loc.append(((cd // win[0], chh // win[1], cw // win[2]), ((cd % win[0]) * win[1] + chh % win[1]) * win[2] + cw % win[2]))
            print("   (window, local index):", sorted(loc)[:40])


if __name__ == "__main__":
    main()
