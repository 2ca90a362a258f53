"""TEST TOOL: c2shape step on the library argv[1]; writes the norm of every GEMM's inputs and outputs (call order) to argv[2]."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import valor_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import valor_b200.kernels as K
from tests.test_gpu_model import run
rows = []
orig = K.gemm
def rec(a, b, **kw):
    r = orig(a, b, **kw)
    outs = r if isinstance(r, tuple) else (r,)
    M, Kd = (a.shape if kw.get("a_kmajor", True) else (a.shape[1], a.shape[0]))
    N = b.shape[0] if kw.get("b_kmajor", True) else b.shape[1]
    rows.append({"i": len(rows), "M": M, "N": N, "K": Kd, "ak": kw.get("a_kmajor", True), "bk": kw.get("b_kmajor", True), "act": kw.get("act", 0),
                 "aux": kw.get("act_aux") is not None, "pre": bool(kw.get("want_preact", False)), "acc": bool(kw.get("accumulate", False)),
                 "a": a.double().norm().item(), "b": b.double().norm().item(), "aux_n": kw["act_aux"].double().norm().item() if kw.get("act_aux") is not None else 0.0,
                 "out": [o.double().norm().item() for o in outs]})
    return r
K.gemm = rec
run("c2shape", torch.bfloat16)
json.dump(rows, open(sys.argv[2], "w"))
print("wrote", len(rows))
