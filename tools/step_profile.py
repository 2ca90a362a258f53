"""Per-kernel breakdown of one training step with warm caches (CUPTI via torch.profiler, eager launches).

ncu serialises and cold-starts every launch, which overstates short or L2-friendly kernels; this tool
gives the in-situ durations used to decide what to optimise next.  Usage:
    python tools/step_profile.py [--batch 32] [--top 45] > gpurun_out/step_profile.txt
"""
import argparse
import os
import re
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--tokens", type=int, default=32)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    from tools import synth   # seeded synthetic weights / batch (the GPU arm never touches oracle/)
    from valor_b200.optim import get_lr_sched
    from valor_b200.pretrain import VALOR, default_opts
    import bench

    dev = torch.device("cuda", 0)
    geom = synth.BASE
    opts = default_opts(swin_depths=geom.swin_depths, ast_layers=geom.ast_layers, bert_layers=geom.bert_layers,
                        num_train_steps=1000)
    model = VALOR.from_pretrained(opts, synth.make_state_dict(geom, seed=0))
    store = model.attach(dtype=torch.bfloat16, device=dev)
    host = synth.make_batch(args.batch, args.frames, args.clips, args.tokens, geom, seed=123)
    tokens = host["txt_tokens"]["bert_tokens"]
    mi, ml = synth.token_masker(tokens, 0.6, seed=1234)
    batch = {"video_pixels": host["video_pixels"].to(dev), "audio_spectrograms": host["audio_spectrograms"].to(dev),
             "txt_tokens": {"bert_tokens": tokens.to(dev)}, "caption_mask": (mi.to(dev), ml.to(dev)), "ids": host["ids"]}

    def step(i):
        store.zero_grad()
        losses = model(batch, bench.TASK, compute_loss=True)
        sum(losses.values()).backward()
        store.set_hyper(get_lr_sched(i + 1, opts), base_lr=opts.learning_rate, betas=tuple(opts.betas),
                        weight_decay=opts.weight_decay)
        store.optimizer_step(max_norm=opts.grad_norm)

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(2 + i)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for ev in prof.events():
        if ev.device_type is None or "cuda" not in str(ev.device_type).lower():
            continue
        name = re.sub(r"\(anonymous namespace\)::|unnamed>::", "", ev.name); name = re.sub(r"\(.*", "", name).replace("void ", "").replace("valor::", "")
        us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        a = agg[name]
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    total = sum(v[1] for v in agg.values())
    n = args.steps
    print(f"kernels/step {sum(v[0] for v in agg.values()) / n:.0f}   device time/step {total / n / 1e3:.2f} ms")
    for name, (cnt, us, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"{us / n / 1e3:9.3f} ms {100 * us / total:5.1f}%  n={cnt / n:6.0f}  avg {us / cnt:9.1f} us  max {mx:9.1f}  {name[:110]}")


if __name__ == "__main__":
    main()
