"""TEST TOOL: run the c2shape step on library A, record every GEMM call (tensors kept), replay each call on library B
and report the calls whose outputs differ beyond bf16 rounding.  usage: gemm_replay_ab.py libA.so libB.so"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import valor_b200._lib as L
libA, libB = (os.path.abspath(p) for p in sys.argv[1:3])
L.LIB_PATH = libA
import valor_b200.kernels as K
from tests.test_gpu_model import run

calls = []
orig = K.gemm


def rec(a, b, **kw):
    pre_out = kw["out"].clone() if kw.get("out") is not None and kw.get("accumulate") else None
    r = orig(a, b, **kw)
    if a.dtype == torch.bfloat16 and len(calls) < 4000:
        outs = r if isinstance(r, tuple) else (r,)
        calls.append((a, b, {k: v for k, v in kw.items()}, pre_out, tuple(o.clone() for o in outs)))
    return r


K.gemm = rec
golden, model, losses = run("c2shape", torch.bfloat16)
K.gemm = orig
torch.cuda.synchronize()
print("recorded", len(calls), "bf16 GEMM calls on", os.path.basename(libA))
# switch library
L._lib = None
L.LIB_PATH = libB
L.load()
bad = 0
for i, (a, b, kw, pre_out, outsA) in enumerate(calls):
    kw = dict(kw)
    if kw.get("out") is not None:
        kw["out"] = pre_out.clone() if pre_out is not None else torch.empty_like(kw["out"])
    if kw.get("bias_grad") is not None:
        kw["bias_grad"] = torch.zeros_like(kw["bias_grad"])
    r = orig(a, b, **kw)
    outsB = r if isinstance(r, tuple) else (r,)
    for j, (oa, ob) in enumerate(zip(outsA, outsB)):
        da, db = oa.float(), ob.float()
        rel = (da - db).norm().item() / (da.norm().item() + 1e-20)
        if kw.get("act", 0) == 1:   # GELU calls: elementwise, in units of the larger magnitude's bf16 ulp
            ulp = torch.maximum(da.abs(), db.abs()) * 2.0 ** -8 + 1e-30
            n_bad = ((da - db).abs() > 2.5 * ulp).sum().item()
            if n_bad:
                idx = ((da - db).abs() / ulp).argmax().item()
                print(f"call {i} out{j}: GELU elementwise mismatch on {n_bad} of {da.numel()} elements; worst {da.flatten()[idx].item():.6g} vs {db.flatten()[idx].item():.6g}", flush=True)
        if rel > 2e-3 or not torch.isfinite(db).all():
            bad += 1
            if bad <= 12:
                M, Kd = (a.shape if kw.get("a_kmajor", True) else (a.shape[1], a.shape[0]))
                N = b.shape[0] if kw.get("b_kmajor", True) else b.shape[1]
                print(f"call {i} out{j}: M={M} N={N} K={Kd} ak={kw.get('a_kmajor', True)} bk={kw.get('b_kmajor', True)} act={kw.get('act', 0)} aux={kw.get('act_aux') is not None} "
                      f"pre={kw.get('want_preact', False)} res={kw.get('residual') is not None} acc={kw.get('accumulate', False)} bias={kw.get('bias') is not None} rel diff {rel:.3e} "
                      f"maxabs {(da - db).abs().max().item():.3e} strides a{tuple(a.stride())} b{tuple(b.stride())}", flush=True)
print("differing outputs:", bad)
