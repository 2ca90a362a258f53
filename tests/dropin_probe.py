"""Child process of tests/test_dropin_reference.py (install() swaps classes process-wide, so it runs isolated):
the UNMODIFIED reference `model/pretrain.py: VALOR` + `optim/misc.py: build_optimizer` + the reference step tail on top
of valor_b200 (CPU: the C ABI is replaced by tests/cpu_backend.py, as in the other host-logic tests)."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cpu_backend  # noqa: E402
import valor_b200.kernels as K  # noqa: E402

for name in dir(cpu_backend):
    obj = getattr(cpu_backend, name)
    if isinstance(obj, types.FunctionType) and not name.startswith("_"):
        setattr(K, name, obj)

from oracle import ref_shim  # noqa: E402
from tools import synth  # noqa: E402
from valor_b200 import dropin  # noqa: E402

ref_shim._install_fake_modules()          # ipdb / tensorboardX / easydict ... (NOT apex: install() below owns it)
for k in [k for k in sys.modules if k == "apex" or k.startswith("apex.")]:
    del sys.modules[k]
swapped = dropin.install(reference_root=ref_shim.REFERENCE_ROOT, dtype=torch.float32)
import apex  # noqa: E402
from apex import amp  # noqa: E402

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_tiny.json")))
cfg = golden["config"]
geom = synth.TINY
sd = synth.make_state_dict(geom, seed=cfg["weight_seed"])
model = ref_shim.build_reference_valor(geom, sd)                   # reference VALOR.__init__ / from_pretrained, our encoder classes
kinds = {n: type(getattr(model, n)).__module__ for n in ("video_encoder", "audio_encoder", "multimodal_encoder", "cls", "audio_embeddings")}
from optim.misc import build_optimizer  # noqa: E402  (reference)
from optim.sched import get_lr_sched  # noqa: E402
from torch.nn.utils import clip_grad_norm_  # noqa: E402
from tests.golden.make_golden import _patch_legacy_overloads  # noqa: E402
_patch_legacy_overloads()
opts = ref_shim.default_opts()
optimizer = build_optimizer(model, opts)
model, optimizer = amp.initialize(model, optimizer, enabled=False, opt_level="O2")     # train_utils.py:222
model.rng.active = False                                             # parity mode (the golden was minted with dropout off)
batch = synth.make_batch(cfg["B"], cfg["F"], cfg["A"], cfg["T"], geom, seed=cfg["batch_seed"])
ti, tl = synth.token_masker(batch["txt_tokens"]["bert_tokens"], 0.6, seed=cfg["mask_seed"])
model.text_masker = ref_shim.FixedMasker(ti, tl)
out = {"swapped": swapped, "kinds": kinds, "steps": []}
named = dict(model.named_parameters())
for step in range(1, 4):
    with ref_shim.cuda_identity():
        loss_dict = model(batch, cfg["task"], compute_loss=True)       # reference VALOR.forward (pretrain.py:125)
    loss = sum(loss_dict.values())
    with amp.scale_loss(loss, optimizer) as scaled:                     # train_utils.py:317-319
        scaled.backward()
    if step == 1:
        out["grad_total_norm"] = model.store.grad.double().pow(2).sum().sqrt().item()
        out["grads"] = {k: named[k].grad.double().norm().item() for k in golden["grads"] if golden["grads"][k] is not None}
    lr_ratio = get_lr_sched(step, opts)
    for g in optimizer.param_groups:
        g["lr"] = g["init_lr"] * lr_ratio
    gn = clip_grad_norm_(amp.master_params(optimizer), opts.grad_norm)  # train_utils.py:359
    optimizer.step()                                                    # reference AdamW
    optimizer.zero_grad()
    out["steps"].append({"losses": {k: v.item() for k, v in loss_dict.items()}, "grad_norm": float(gn)})
out["params"] = {k: named[k].detach().double().norm().item() for k in golden["trajectory"]["params"] if k in named}
print("DROPIN_RESULT " + json.dumps(out))
