"""Micro-benchmark of the tcgen05 GEMM (valor_gemm) on the hot shapes of the VALOR-base step.
Prints one JSON line per shape: achieved TFLOP/s vs the measured cuBLAS peak (MEASURED_PEAKS.json),
and torch.matmul (cuBLAS) on the same shape for reference."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valor_b200 import kernels as K  # noqa: E402

SHAPES = [  # (name, M, N, K, form)
    ("swin1.qkv", 802816, 384, 128, "nt"), ("swin1.fc1", 802816, 512, 128, "nt"), ("swin1.fc2", 802816, 128, 512, "nt"),
    ("swin2.qkv", 200704, 768, 256, "nt"), ("swin2.fc1", 200704, 1024, 256, "nt"),
    ("swin3.qkv", 50176, 1536, 512, "nt"), ("swin3.proj", 50176, 512, 512, "nt"), ("swin3.fc1", 50176, 2048, 512, "nt"),
    ("swin3.fc2", 50176, 512, 2048, "nt"), ("swin4.fc1", 12544, 4096, 1024, "nt"),
    ("ast.qkv", 8256, 2304, 768, "nt"), ("ast.fc1", 8256, 3072, 768, "nt"),
    ("bert.qkv", 3072, 2304, 768, "nt"), ("bert.kv", 20800, 1536, 768, "nt"), ("bert.fc1", 3072, 3072, 768, "nt"),
    ("mlm.logits", 3072, 30522, 768, "nt"),
    ("swin3.fc1.dgrad", 50176, 512, 2048, "nn"), ("swin1.fc1.dgrad", 802816, 128, 512, "nn"),
    ("swin3.fc1.wgrad", 2048, 512, 50176, "tn"), ("swin1.qkv.wgrad", 384, 128, 802816, "tn"),
    ("bert.fc1.wgrad", 3072, 768, 3072, "tn"), ("big", 8192, 8192, 8192, "nt"),
    # fused epilogues: fc1 forward (bias + erf-GELU, pre-activation kept), fc2 dgrad (* GELU'(pre)), proj/fc2 (+ residual)
    ("epi.swin1.fc1.gelu_pre", 802816, 512, 128, "nt"), ("epi.swin3.fc1.gelu_pre", 50176, 2048, 512, "nt"),
    ("epi.swin1.fc2.dgrad.gelu_aux", 802816, 512, 128, "nn"), ("epi.swin3.fc2.dgrad.gelu_aux", 50176, 2048, 512, "nn"),
    ("epi.swin1.fc2.res", 802816, 128, 512, "nt"), ("epi.swin3.fc2.res", 50176, 512, 2048, "nt"),
]


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


WGRAD = [  # (name, rows, out features, in features): Linear backward of the VALOR-base step
    ("swin1.qkv", 802816, 384, 128), ("swin1.proj", 802816, 128, 128), ("swin1.fc1", 802816, 512, 128), ("swin1.fc2", 802816, 128, 512),
    ("swin2.qkv", 200704, 768, 256), ("swin2.fc2", 200704, 256, 1024),
    ("swin3.qkv", 50176, 1536, 512), ("swin3.proj", 50176, 512, 512), ("swin3.fc1", 50176, 2048, 512), ("swin3.fc2", 50176, 512, 2048),
    ("swin4.fc1", 12544, 4096, 1024), ("ast.qkv", 8256, 2304, 768), ("ast.fc2", 8256, 768, 3072),
    ("bert.qkv", 3072, 2304, 768), ("bert.proj", 3072, 768, 768), ("bert.fc1", 3072, 3072, 768), ("bert.kv", 20800, 1536, 768),
]


def wgrad_ab():
    """Linear backward: weight-gradient launch + column-sum launch vs the single launch with the bias gradient fused."""
    for name, rows, nout, nin in WGRAD:
        dy = torch.randn(rows, nout, device="cuda", dtype=torch.bfloat16)
        x = torch.randn(rows, nin, device="cuda", dtype=torch.bfloat16)
        dw = torch.zeros(nout, nin, device="cuda")
        db = torch.zeros(nout, device="cuda")

        def two():
            K.gemm(dy, x, a_kmajor=False, b_kmajor=False, out=dw, accumulate=True)
            K.colsum(dy, db)

        ms_g = bench(lambda: K.gemm(dy, x, a_kmajor=False, b_kmajor=False, out=dw, accumulate=True))
        ms_two = bench(two)
        ms_one = bench(lambda: K.gemm(dy, x, a_kmajor=False, b_kmajor=False, out=dw, accumulate=True, bias_grad=db))
        print(json.dumps({"linear": name, "rows": rows, "out": nout, "in": nin, "wgrad_ms": round(ms_g, 4),
                          "wgrad_plus_colsum_ms": round(ms_two, 4), "fused_ms": round(ms_one, 4)}), flush=True)


def main():
    if "--wgrad-ab" in sys.argv:
        return wgrad_ab()
    peak = 1400.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("bf16_tflops", peak)
    only = sys.argv[1:]
    for name, M, N, Kd, form in SHAPES:
        if only and not any(o in name for o in only):
            continue
        a_k, b_k = {"nt": (True, True), "nn": (True, False), "tn": (False, False)}[form]
        a = torch.randn((M, Kd) if a_k else (Kd, M), device="cuda", dtype=torch.bfloat16)
        b = torch.randn((N, Kd) if b_k else (Kd, N), device="cuda", dtype=torch.bfloat16)
        acc = form == "tn"
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if acc else torch.bfloat16)
        kw = {}
        if "gelu_pre" in name:
            kw = dict(bias=torch.randn(N, device="cuda"), act=K.ACT_GELU, want_preact=True)
        elif "gelu_aux" in name:
            kw = dict(act=K.ACT_GELU, act_aux=torch.randn(M, N, device="cuda", dtype=torch.bfloat16))
        elif name.endswith(".res"):
            kw = dict(bias=torch.randn(N, device="cuda"), residual=torch.randn(M, N, device="cuda", dtype=torch.bfloat16))
        ms = bench(lambda: K.gemm(a, b, a_kmajor=a_k, b_kmajor=b_k, out=out, accumulate=acc, **kw))
        A = a if a_k else a.t()
        Bt = b.t() if b_k else b
        ms_cublas = bench(lambda: torch.matmul(A, Bt))
        fl = 2.0 * M * N * Kd
        print(json.dumps({"shape": name, "M": M, "N": N, "K": Kd, "form": form, "ms": round(ms, 4),
                          "tflops": round(fl / ms / 1e9, 1), "frac_of_peak": round(fl / ms / 1e9 / peak, 3),
                          "cublas_ms": round(ms_cublas, 4), "cublas_tflops": round(fl / ms_cublas / 1e9, 1),
                          "gbytes_per_s": round((a.numel() + b.numel() + out.numel()) * 2 / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
