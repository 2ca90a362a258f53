"""TEST TOOL: per-block cycle stamps of one CTA of the window backward (experiment build libvalor_exp8.so, see
build_exp.sh).  Prints, per (k-tile, q-tile) block, where the element warps and the MMA thread spend their cycles."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import valor_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.join(HERE, "libvalor_exp8.so")
import torch  # noqa: E402
from valor_b200 import kernels as K  # noqa: E402


def main():
    B, D, H, W, heads, hd = 4, 8, 56, 56, 4, 32
    shift = (0, 3, 3) if "--shift" in sys.argv else (0, 0, 0)
    C = heads * hd
    tokens = B * D * H * W
    qkv = torch.randn(tokens, 3 * C, device="cuda", dtype=torch.bfloat16)
    table = torch.randn(15 * 13 * 13, heads, device="cuda") * 0.5
    do = torch.randn(tokens, C, device="cuda", dtype=torch.bfloat16)
    geom = ((B, D, H, W), (8, 7, 7), shift, (8, 7, 7), heads, hd, hd ** -0.5)
    o, lse = K.window_attn_fwd(qkv, table, *geom)
    dt = None if "--nodtab" in sys.argv else torch.zeros_like(table)
    for _ in range(3):
        K.window_attn_bwd(qkv, o, do, lse, table, dt, *geom)
    torch.cuda.synchronize()
    lib = ctypes.CDLL(L.LIB_PATH)
    n = 64 * 32
    buf = (ctypes.c_longlong * n)()
    rc = lib.valor_exp_trace(buf, n)
    assert rc == 0, rc
    t = [[buf[b * 32 + s] for s in range(32)] for b in range(16)]
    t0 = t[0][16]
    names = ["e.pre", "e.sfull", "e.math", "e.pdfree", "e.pdfull", "e.bar5", "e.flush"]
    print("cycles relative to the MMA thread's first stamp; element stamps: half 0 warp (half 1 warp)")
    prev_end = None
    for b in range(16):
        r = t[b]
        el = " ".join(f"{names[i]}={r[i] - t0}({r[8 + i] - t0})" for i in range(7))
        print(f"      warp0: sfull->ld0 {r[20] - r[1]}  math0 {r[21] - r[20]}  ld1 {r[22] - r[21]}  math1 {r[23] - r[22]}")
        print(f"b={b:2d} kt={b // 4} qt={b % 4} | mma: top={r[16] - t0} s_issued={r[17] - t0} pdfull={r[18] - t0} kv_issued={r[19] - t0} | {el}")
    print("block period (MMA thread top-of-loop):", [t[b + 1][16] - t[b][16] for b in range(15)])


if __name__ == "__main__":
    main()
