"""SURVEY §8b drop-in: after `valor_b200.dropin.install()` the unmodified reference model (`model/pretrain.py: VALOR`),
optimizer (`optim/misc.py`, `optim/adamw.py`) and step tail (`train_utils.py:306-363` with the `apex.amp` calls it makes)
run on this library and reproduce the goldens minted from the stock reference.  Needs /root/reference."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="needs /root/reference (build container)")
def test_unmodified_reference_model_and_step_run_on_the_library():
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_probe.py")], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("DROPIN_RESULT ")]
    assert line, r.stderr[-2000:]
    out = json.loads(line[-1][len("DROPIN_RESULT "):])
    golden = json.load(open(os.path.join(HERE, "golden", "golden_tiny.json")))
    # the reference constructed OUR encoder classes by name
    assert all(v.startswith("valor_b200.") for v in out["kinds"].values()), out["kinds"]
    assert {"model.videoswin.SwinTransformer3D", "model.bert.BertModel", "model.modeling.BERTPredictionHead"} <= set(out["swapped"])
    for k, v in golden["losses"].items():
        assert abs(out["steps"][0]["losses"][k] - v) <= 1e-4 * abs(v), (k, out["steps"][0]["losses"][k], v)
    assert abs(out["grad_total_norm"] - golden["grad_total_norm"]) <= 5e-4 * golden["grad_total_norm"]
    for k, ref in golden["grads"].items():
        if ref is not None:
            assert abs(out["grads"][k] - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-9, k
    traj = golden["trajectory"]
    for i, rec in enumerate(traj["steps"]):
        assert abs(out["steps"][i]["grad_norm"] - rec["grad_norm"]) <= 5e-4 * rec["grad_norm"]
        if rec["losses_before"]:
            for k, v in rec["losses_before"].items():
                assert abs(out["steps"][i]["losses"][k] - v) <= 2e-4 * abs(v)
    for k, ref in traj["params"].items():
        assert abs(out["params"][k] - ref["norm"]) <= 1e-5 * ref["norm"] + 1e-9, k
