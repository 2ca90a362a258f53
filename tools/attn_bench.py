"""Micro-benchmark of the window / MHA attention kernels at the VALOR-base shapes."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valor_b200 import kernels as K  # noqa: E402


def bench(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def window(B, grid_hw, heads, shift, name, D=8, backend=0):
    H, W = grid_hw, grid_hw   # C2: 8 frames -> 8 temporal tokens (frames are embedded in pairs with themselves), window (8,7,7)
    hd, C = 32, heads * 32
    tokens = B * D * H * W
    qkv = torch.randn(tokens, 3 * C, device="cuda", dtype=torch.bfloat16)
    table = torch.randn(15 * 13 * 13, heads, device="cuda") * 0.5
    do = torch.randn(tokens, C, device="cuda", dtype=torch.bfloat16)
    geom = ((B, D, H, W), (min(D, 8), 7, 7), shift, (8, 7, 7), heads, hd, hd ** -0.5)
    N = min(D, 8) * 49
    o, lse = K.window_attn_fwd(qkv, table, *geom, backend=backend)
    dt = None if "--nodtab" in sys.argv else torch.zeros_like(table)
    ms_f = bench(lambda: K.window_attn_fwd(qkv, table, *geom, backend=backend))
    ms_b = bench(lambda: K.window_attn_bwd(qkv, o, do, lse, table, dt, *geom, backend=backend))
    nprob = B * (D // min(D, 8)) * (H // 7) * (W // 7) * heads
    fl = 4.0 * N * N * hd * nprob
    print(json.dumps({"kernel": name, "backend": {0: "auto (mma.sync fwd + tcgen05 bwd)", 1: "tcgen05", 3: "mma.sync (round 1)"}.get(backend, backend), "problems": nprob, "fwd_ms": round(ms_f, 3), "bwd_ms": round(ms_b, 3),
                      "fwd_tflops": round(fl / ms_f / 1e9, 1), "bwd_tflops": round(2.5 * fl / ms_b / 1e9, 1),
                      "elems_per_ns_fwd": round(N * N * nprob / ms_f / 1e6, 1)}), flush=True)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["s1", "s1s", "s3"]
    for be in ((3, 1, 0) if "--both" in sys.argv else (0,)):
        if "s1" in which:
            window(32, 56, 4, (0, 0, 0), "swin stage1 (no shift)", backend=be)
        if "s1s" in which:
            window(32, 56, 4, (0, 3, 3), "swin stage1 (shifted)", backend=be)
        if "s3" in which:
            window(32, 14, 16, (0, 3, 3), "swin stage3 (shifted)", backend=be)
        if "d4" in which:
            window(32, 56, 4, (0, 3, 3), "swin stage1, 4 temporal tokens (196-token windows)", D=4, backend=be)


if __name__ == "__main__":
    main()
