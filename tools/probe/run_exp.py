"""TEST TOOL: times the experiment builds of the window backward (see build_exp.sh) at the stage-1 shape."""
import os, sys, subprocess, json
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if len(sys.argv) > 1 and sys.argv[1].endswith(".so"):
    import valor_b200._lib as L
    L.LIB_PATH = sys.argv[1]
    sys.argv = [sys.argv[0], "s1"] + sys.argv[2:]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import attn_bench
    attn_bench.main()
else:
    extra = [a for a in sys.argv[1:]]
    for v, what in ((0, "product"), (3, "no element math, no dK/dV/dQ MMAs"), (7, "prologue + teardown only")):
        lib = os.path.join(ROOT, "valor_b200", "csrc", "libvalor_b200.so") if v == 0 else os.path.join(HERE, f"libvalor_exp{v}.so")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), lib] + extra, capture_output=True, text=True, cwd=ROOT, env={**os.environ, "PYTHONPATH": ROOT})
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(what, "->", line[-1] if line else r.stderr[-300:], flush=True)
