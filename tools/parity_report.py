"""Parity report (GPU box): deviation of the CUDA path from the live-reference goldens, every config x dtype.
   python tools/parity_report.py  -> gpurun_out/parity_report.json (copy the final one into profiles/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_host_logic import build, masked_logits  # noqa: E402


def report(name, dtype):
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", f"golden_{name}.json")))
    model, batch = build(golden["config"], dtype=dtype, device="cuda")
    model.debug_capture = {}
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    model.store.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    out = {"config": name, "dtype": str(dtype).split(".")[-1], "loss_rel_err": {}, "grad_norm_rel_err": {}}
    for k, v in golden["losses"].items():
        out["loss_rel_err"][k] = abs(losses[k].item() - v) / abs(v)
    tot, tot_ref = sum(v.item() for v in losses.values()), sum(golden["losses"].values())
    out["loss_rel_err"]["total"] = abs(tot - tot_ref) / abs(tot_ref)
    total = model.store.grad.double().pow(2).sum().sqrt().item()
    out["grad_total_norm_rel_err"] = abs(total - golden["grad_total_norm"]) / golden["grad_total_norm"]
    named = dict(model.named_parameters())
    for k, ref in golden["grads"].items():
        if ref is None:
            continue
        out["grad_norm_rel_err"][k] = abs(named[k].main_grad.double().norm().item() - ref["norm"]) / (ref["norm"] + 1e-12)
    out["grad_norm_rel_err_max"] = max(out["grad_norm_rel_err"].values())
    got = masked_logits(model, golden)
    worst = 0.0
    for nm, ref in golden["logits"].items():
        rows = len(ref["lse"])
        worst = max(worst, (got[nm][:rows, :8].cpu() - torch.tensor(ref["head"])).abs().max().item())
        worst = max(worst, (torch.logsumexp(got[nm][:rows], -1).cpu() - torch.tensor(ref["lse"])).abs().max().item())
    out["logits_max_abs_err"] = worst
    return out


def report_torch_eager_bf16(name):
    """the same deviations for stock torch eager in bf16 autocast (the oracle's restatement of the reference modules on
    the GPU): separates what bf16 arithmetic costs in ANY implementation from what valor_b200's kernels add"""
    from oracle import valor_oracle as vo
    from tools import synth
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", f"golden_{name}.json")))
    cfg = golden["config"]
    geom = {"tiny": synth.TINY, "c1": synth.BASE}[cfg["geom"]]
    sd = synth.make_state_dict(geom, seed=cfg["weight_seed"], include_buffers=False)
    params = {k: v.cuda().requires_grad_(True) for k, v in sd.items()
              if not k.startswith("txt_encoder.") and k != "cls.decoder.weight"}
    full = dict(params)
    for k in sd:
        if k.startswith("txt_encoder."):
            full[k] = params["multimodal_encoder." + k[len("txt_encoder."):]]
    full["cls.decoder.weight"] = params["multimodal_encoder.embeddings.word_embeddings.weight"]
    b = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
    ti, tl = synth.token_masker(b["txt_tokens"]["bert_tokens"], 0.6, seed=cfg["mask_seed"])
    batch = {"video_pixels": b["video_pixels"].cuda(), "audio_spectrograms": b["audio_spectrograms"].cuda(),
             "txt_tokens": {"bert_tokens": b["txt_tokens"]["bert_tokens"].cuda()}}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        losses = vo.forward_pt(batch, full, geom, ti.cuda(), tl.cuda(), task=cfg["task"])
    sum(losses.values()).float().backward()
    out = {"config": name, "dtype": "torch-eager bf16 autocast (oracle port)", "loss_rel_err": {}, "grad_norm_rel_err": {}}
    for k, v in golden["losses"].items():
        out["loss_rel_err"][k] = abs(losses[k].item() - v) / abs(v)
    tot = sum(p.grad.double().pow(2).sum().item() for p in params.values() if p.grad is not None) ** 0.5
    out["grad_total_norm_rel_err"] = abs(tot - golden["grad_total_norm"]) / golden["grad_total_norm"]
    for k, ref in golden["grads"].items():
        if ref is None or params[k].grad is None:
            continue
        out["grad_norm_rel_err"][k] = abs(params[k].grad.double().norm().item() - ref["norm"]) / (ref["norm"] + 1e-12)
    out["grad_norm_rel_err_max"] = max(out["grad_norm_rel_err"].values())
    return out


if __name__ == "__main__":
    rows = []
    for name in ("tiny", "c1", "c2shape"):
        for dtype in (torch.float32, torch.bfloat16):
            r = report(name, dtype)
            rows.append(r)
            print(json.dumps({k: v for k, v in r.items() if k != "grad_norm_rel_err"}), flush=True)
    for name in ("tiny", "c1", "c2shape"):
        try:
            r = report_torch_eager_bf16(name)
            rows.append(r)
            print(json.dumps({k: v for k, v in r.items() if k != "grad_norm_rel_err"}), flush=True)
        except Exception as ex:  # pragma: no cover
            print("torch eager comparison failed:", type(ex).__name__, ex, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
