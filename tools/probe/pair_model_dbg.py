"""TEST TOOL: the c2shape per-parameter gradient check against a given build of the library (argv[1])."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import valor_b200._lib as L
if len(sys.argv) > 1:
    L.LIB_PATH = os.path.abspath(sys.argv[1])
import valor_b200.kernels as K
from tests.test_gpu_model import run
for name in sys.argv[2:] or ["c2shape"]:
    golden, model, losses = run(name, torch.bfloat16)
    named = dict(model.named_parameters())
    total = golden["grad_total_norm"]
    rows = []
    for k, ref in golden["grads"].items():
        if ref is None: continue
        g = named[k].main_grad
        dev = abs(g.double().norm().item() - ref["norm"])
        rows.append((dev / (ref["norm"] + 1e-12), k, g.double().norm().item(), ref["norm"]))
    rows.sort(reverse=True)
    print(os.path.basename(L.LIB_PATH), name, {k: round(float(v), 5) for k, v in losses.items()}, "gold", {k: round(v, 5) for k, v in golden["losses"].items()},
          "total", round(model.store.grad.double().norm().item(), 4), round(total, 4))
    for w in rows[:8]:
        print("   rel %.4f  %-70s got %.5f ref %.5f" % w)
