"""Checkpoint interop with the reference's on-disk format.

  model_step_N.pt      = CPU state_dict with the reference's keys (utils/save.py:43-46), including the
                         `txt_encoder.*` duplicates of the shared BERT (modeling.py:685-691);
  optimizer_step_N.pt  = the torch-optimizer state_dict of optim/misc.py:build_optimizer's 10 parameter
                         groups (utils/save.py:58-64): {'state': {index: {'step','exp_avg','exp_avg_sq'}},
                         'param_groups': [...]}, parameters indexed in group order.

`prepare_checkpoint` applies the key / shape adaptations of train_utils.load_from_pretrained_dir (:129-146):
strip DDP's `module.` prefix and extend the learned frame embeddings past the pretraining frame count by
repeating the last trained slot.  Nothing here touches the GPU kernels: the arenas of params.ParamStore are the
storage, this module only maps them to and from the reference's file layout.
"""
import os

import torch

from .params import is_no_decay


def prepare_checkpoint(checkpoint, video_sample_num=None, audio_sample_num=None):
    """train_utils.py:129-146 (CLIP position-embedding interpolation :148-166 belongs to the CLIP tower)."""
    ck = {k.replace("module.", ""): v for k, v in checkpoint.items()}
    if "video_frame_embedding" in ck and video_sample_num:
        e = ck["video_frame_embedding"].clone()
        e[:, video_sample_num:] = e[:, video_sample_num - 1].clone().unsqueeze(1)
        ck["video_frame_embedding"] = e
    if "audio_frame_embedding" in ck and audio_sample_num:
        e = ck["audio_frame_embedding"].clone()
        e[:, audio_sample_num:] = e[:, audio_sample_num - 1].clone().unsqueeze(1)
        ck["audio_frame_embedding"] = e
    return ck


def reference_param_groups(model, opts):
    """optim/misc.py:13-77 for a VideoSwin/BERT model (no `clip` / `multimodal_encoder.decoder` / new_params names):
    group 0 = decayed, group 1 = no-decay, groups 2-9 empty.  Returns (groups, ordered (name, param) list)."""
    decay, no_decay = [], []
    for k, v in model.named_parameters():
        (no_decay if is_no_decay(k) else decay).append((k, v))
    lr = opts.learning_rate
    specs = [(decay, opts.weight_decay, lr), (no_decay, 0.0, lr)] + [([], wd, l) for wd, l in (
        (opts.weight_decay, getattr(opts, "new_lr", 0.0)), (0.0, getattr(opts, "new_lr", 0.0)),
        (opts.weight_decay, getattr(opts, "clip_lr", 5e-7)), (0.0, getattr(opts, "clip_lr", 5e-7)),
        (opts.weight_decay, getattr(opts, "clip_lr_text", 5e-7)), (0.0, getattr(opts, "clip_lr_text", 5e-7)),
        (opts.weight_decay, lr), (0.0, lr))]
    return specs, decay + no_decay


def optimizer_state_dict(model, opts, lr_ratio=1.0):
    """The store's AdamW state in the reference optimizer's state_dict layout."""
    st = model.store
    specs, ordered = reference_param_groups(model, opts)
    state, groups, idx = {}, [], 0
    steps = dict(zip(st.names, st.param_steps))
    for plist, wd, lr in specs:
        ids = []
        for name, p in plist:
            off, k = st.offsets[name]
            if steps[name] > 0:
                state[idx] = {"step": steps[name], "exp_avg": st.exp_avg[off:off + k].view(p.shape).detach().cpu().clone(),
                              "exp_avg_sq": st.exp_avg_sq[off:off + k].view(p.shape).detach().cpu().clone()}
            ids.append(idx)
            idx += 1
        groups.append({"weight_decay": wd, "lr": lr * lr_ratio, "init_lr": lr, "betas": tuple(opts.betas), "eps": 1e-6,
                       "correct_bias": True, "params": ids})
    return {"state": state, "param_groups": groups}


def load_optimizer_state(model, opts, opt_state):
    st = model.store
    _, ordered = reference_param_groups(model, opts)
    index = {n: i for i, n in enumerate(st.names)}
    for i, (name, p) in enumerate(ordered):
        s = opt_state["state"].get(i)
        off, k = st.offsets[name]
        if s is None:
            st.param_steps[index[name]] = 0
            st.exp_avg[off:off + k].zero_()
            st.exp_avg_sq[off:off + k].zero_()
            continue
        st.param_steps[index[name]] = int(s["step"])
        st.exp_avg[off:off + k].copy_(s["exp_avg"].reshape(-1))
        st.exp_avg_sq[off:off + k].copy_(s["exp_avg_sq"].reshape(-1))
    st.step = max(st.param_steps) if st.param_steps else 0
    st._runs = None


def save_checkpoint(model, opts, out_dir, step, with_optimizer=True, lr_ratio=1.0):
    """utils/save.py:32-64 (model_step_N.pt [+ optimizer_step_N.pt]); returns the two paths."""
    os.makedirs(out_dir, exist_ok=True)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    mp = os.path.join(out_dir, f"model_step_{step}.pt")
    torch.save(sd, mp)
    op = None
    if with_optimizer and model.store is not None:
        op = os.path.join(out_dir, f"optimizer_step_{step}.pt")
        torch.save(optimizer_state_dict(model, opts, lr_ratio), op)
    return mp, op


def load_checkpoint(model, path, video_sample_num=None, audio_sample_num=None):
    """Load a reference-format model file into an attached (or not yet attached) model; with a store the fp32 master
    arena is the destination and the bf16 working copy is refreshed."""
    ck = prepare_checkpoint(torch.load(path, map_location="cpu"), video_sample_num, audio_sample_num)
    own = model.state_dict()
    missing = [k for k in own if k not in ck and "relative_position_index" not in k]
    unexpected = [k for k in ck if k not in own]
    with torch.no_grad():
        for k, v in ck.items():
            if k in own:
                own[k].copy_(v.to(own[k].dtype))   # state_dict tensors are views into the master arena
    if model.store is not None and model.store.master.is_cuda:
        model.store.refresh_lp()
    return missing, unexpected
