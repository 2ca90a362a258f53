"""TEST TOOL: sweeps tcgen05 operand-descriptor candidates on a B200 (one subprocess per case so a faulting
candidate cannot poison the others).   python tools/probe/run_probe.py [case ...]  -> gpurun_out/umma_probe.json"""
import ctypes
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def desc(lbo, sbo, layout):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(N, a_mn, b_mn, M=128):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def map_sw64(rows):
    """[rows, 32] bf16 image (4 chunks per row) -> the attention tile layout (window_attn.cu tile_off)"""
    m = []
    for r in range(rows):
        for c in range(4):
            m.append(r * 64 + ((c ^ ((r >> 1) & 3)) << 4))
    return m


def map_sw128_panels(rows, cols):
    """[rows, cols] bf16 image, cols % 64 == 0 -> 64-column panels of rows x 128 B, 128B swizzle"""
    m = []
    for r in range(rows):
        for cc in range(cols // 8):
            p, ch = cc // 8, cc % 8
            m.append(p * rows * 128 + r * 128 + ((ch ^ (r & 7)) << 4))
    return m


CASES = {}
for lbo in (0, 16):
    # S = Q.K^T : A [128,32] and B [128,32] both K-major, 64-byte rows, SWIZZLE_64B
    CASES[f"kmajor_sw64_lbo{lbo}"] = dict(kind="k64", lbo=lbo)
for lbo in (0, 512, 64, 1024):
    # O = P.V : A = P [128 q, 64 keys] K-major SW128 panel; B = V [64 keys, 32] stored [key][ch] -> MN-major SW64
    CASES[f"b_mnmajor_sw64_lbo{lbo}"] = dict(kind="pv", lbo=lbo)
for lbo in (0, 512):
    # dV = P^T.dO : A = P^T from the [q][keys] panels (MN-major SW128), B = dO [128 q, 32] MN-major SW64
    CASES[f"a_mnmajor_sw128_b_mn64_lbo{lbo}"] = dict(kind="ptdo", lbo=lbo)
for lbo in (0, 512):
    # TS form: A (bf16 pairs) in tensor memory, B = V MN-major SW64
    CASES[f"ts_a_tmem_lbo{lbo}"] = dict(kind="ts", lbo=lbo)


def run_case(name):
    c = CASES[name]
    lib = ctypes.CDLL(os.path.join(HERE, "libumma_probe.so"))
    lib.umma_probe.restype = ctypes.c_int
    lib.umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                               ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    g = torch.Generator().manual_seed(1)
    dev = "cuda"
    tw = None
    a_tmem, tcols = 0, 0
    if c["kind"] == "k64":
        A = torch.randn(128, 32, generator=g).bfloat16()
        B = torch.randn(128, 32, generator=g).bfloat16()
        ref = A.float() @ B.float().t()
        amap, bmap = map_sw64(128), map_sw64(128)
        ad = bd = desc(c["lbo"], 512, 4)
        args = dict(a_step=32, b_step=32, ksteps=2, idesc=idesc(128, 0, 0), N=128)
        a_img, b_img = A, B
    elif c["kind"] == "pv":
        A = torch.randn(128, 64, generator=g).bfloat16()     # P  [q, keys]
        V = torch.randn(64, 32, generator=g).bfloat16()      # V  [keys, ch]
        ref = A.float() @ V.float()
        amap, bmap = map_sw128_panels(128, 64), map_sw64(64)
        ad, bd = desc(0, 1024, 2), desc(c["lbo"], 512, 4)
        args = dict(a_step=32, b_step=16 * 64, ksteps=4, idesc=idesc(32, 0, 1), N=32)
        a_img, b_img = A, V
    elif c["kind"] == "ptdo":
        Pm = torch.randn(128, 128, generator=g).bfloat16()   # P [q, keys]  (A = P^T: M = keys, K = q)
        dO = torch.randn(128, 32, generator=g).bfloat16()    # dO [q, ch]
        ref = Pm.float().t() @ dO.float()
        amap, bmap = map_sw128_panels(128, 128), map_sw64(128)
        ad, bd = desc(128 * 128, 1024, 2), desc(c["lbo"], 512, 4)   # A: LBO = next 64-key panel, SBO = 8 q-rows
        args = dict(a_step=16 * 128, b_step=16 * 64, ksteps=8, idesc=idesc(32, 1, 1), N=32)
        a_img, b_img = Pm, dO
    else:
        A = torch.randn(128, 128, generator=g).bfloat16()    # P [q, keys] in tensor memory
        V = torch.randn(128, 32, generator=g).bfloat16()
        ref = A.float() @ V.float()
        amap, bmap = [-1], map_sw64(128)
        ad, bd = 0, desc(c["lbo"], 512, 4)
        args = dict(a_step=0, b_step=16 * 64, ksteps=8, idesc=idesc(32, 0, 1), N=32)
        a_img, b_img = torch.zeros(8, dtype=torch.bfloat16), V
        tw = A.view(torch.int32).contiguous().to(dev)        # [128, 64] words: (k even | k odd << 16)
        a_tmem, tcols = 1, 64
    a_d, b_d = a_img.contiguous().to(dev), b_img.contiguous().to(dev)
    am = torch.tensor(amap, dtype=torch.int32, device=dev)
    bm = torch.tensor(bmap, dtype=torch.int32, device=dev)
    D = torch.zeros(128, args["N"], device=dev)
    rc = lib.umma_probe(a_d.data_ptr(), am.data_ptr(), len(amap) if amap[0] >= 0 else 0, b_d.data_ptr(), bm.data_ptr(), len(bmap),
                        ad, bd, args["a_step"], args["b_step"], args["ksteps"], args["idesc"], args["N"], a_tmem,
                        tw.data_ptr() if tw is not None else None, tcols, D.data_ptr())
    err = (D.cpu() - ref).abs().max().item() if rc == 0 else float("nan")
    print(json.dumps({"case": name, "rc": rc, "max_abs_err": err, "ref_scale": ref.abs().max().item()}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    out = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out.append(json.loads(line[-1]) if line else {"case": n, "rc": r.returncode, "stderr": r.stderr[-300:]})
        except subprocess.TimeoutExpired:
            out.append({"case": n, "rc": "timeout"})
        print(out[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w"), indent=1)
