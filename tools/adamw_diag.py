"""GPU diagnostic: one optimizer step of the CUDA path vs optim/adamw.py semantics recomputed on the host from the
CUDA path's own gradients.  Prints the parameters whose update deviates."""
import json, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_host_logic import build
from valor_b200.optim import get_lr_sched
from valor_b200.pretrain import default_opts
from valor_b200.params import is_no_decay

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_tiny.json")))
model, batch = build(golden["config"], dtype=torch.float32, device="cuda")
opts = default_opts(num_train_steps=1000)
st = model.store
for step in range(1, 3):
    losses = model(batch, golden["config"]["task"], compute_loss=True)
    st.zero_grad()
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    p0, g0 = st.master.clone().cpu(), st.grad.clone().cpu()
    m0, v0 = st.exp_avg.clone().cpu(), st.exp_avg_sq.clone().cpu()
    st.set_hyper(get_lr_sched(step, opts), base_lr=opts.learning_rate, betas=tuple(opts.betas), weight_decay=opts.weight_decay)
    st.optimizer_step(max_norm=opts.grad_norm)
    torch.cuda.synchronize()
    p1 = st.master.clone().cpu()
    print("step", step, "runs", st._runs, "hyper", st.hyper_table[:len(st._runs)].cpu().tolist(), "norm", st.norm.cpu().tolist())
    coef = min(1.0, 5.0 / (g0.double().pow(2).sum().sqrt().item() + 1e-6))
    lr = 1e-4 * get_lr_sched(step, opts)
    worst = []
    for name in st.names:
        off, k = st.offsets[name]
        if name in st._unused:
            d = (p1[off:off + k] - p0[off:off + k]).abs().max().item()
            if d > 0:
                worst.append((d, name, "UNUSED MOVED"))
            continue
        g = g0[off:off + k] * coef
        m = m0[off:off + k] * 0.9 + 0.1 * g
        v = v0[off:off + k] * 0.98 + 0.02 * g * g
        ss = lr * math.sqrt(1 - 0.98 ** step) / (1 - 0.9 ** step)
        p = p0[off:off + k] - ss * m / (v.sqrt() + 1e-6)
        if not is_no_decay(name):
            p = p - lr * 0.01 * p
        d = (p1[off:off + k] - p).abs().max().item()
        mv = (p - p0[off:off + k]).abs().max().item()
        worst.append((d, name, f"expected max move {mv:.3e}"))
    worst.sort(reverse=True)
    for w in worst[:8]:
        print("   ", w)

w = dict(model.named_parameters())["multimodal_encoder.embeddings.word_embeddings.weight"]
print("word emb norm: gpu fp32", w.data.norm().item(), "gpu f64", w.data.double().norm().item(), "cpu", w.data.cpu().norm().item(),
      "cpu f64", w.data.cpu().double().norm().item(), "sum sq gpu", (w.data * w.data).sum().item())
