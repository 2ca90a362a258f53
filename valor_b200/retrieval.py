"""Retrieval scoring and recall (BASELINE configs[4]) — mirror of
`test.py: validate_ret (:249-411)` and `compute_metric_ret (:714-775)` for the fine-grained contrastive head.

The reference gathers the features of the whole evaluation set, then on rank 0 evaluates
`compute_fine_matrix` (pretrain.py:178-211; sliced in 100-row chunks above 1200 candidates because the
[Nt, Nv, T, V] einsum intermediate does not fit) and ranks with a full sort + Python `list.index` per query.
Here the similarity is ONE split-bf16 tcgen05 GEMM [Nt*T, Nv*V] (fp32-grade dot products) + the max/max reduction
kernel of the training path (functional.FineSimFn), and the rank of the ground truth is counted on the device
(valor_retrieval_rank): nothing leaves the GPU but the five summary numbers.
"""
import torch

from . import functional as Fn
from . import kernels as K


def _fine_scores(model, feat_t, feat_v, feat_a, txt_tokens, group):
    """group in {"tva", "tv", "ta"} (test.py:303-345)."""
    Nt, T, D = feat_t.shape
    parts = {"tva": (feat_v, feat_a), "tv": (feat_v, None), "ta": (None, feat_a)}[group]
    fv, fa = parts
    Nv = (fv if fv is not None else fa).shape[0]
    nV = fv.shape[1] if fv is not None else 0
    nA = fa.shape[1] if fa is not None else 0
    feat_t = feat_t.float().contiguous()
    w_t = model._fine_weight(feat_t.reshape(Nt * T, D), "text").view(Nt, T)
    dev = feat_t.device
    w_v = model._fine_weight(fv.float().reshape(Nv * nV, D), "video").view(Nv, nV) if fv is not None else torch.zeros(Nv, 0, device=dev)
    w_a = model._fine_weight(fa.float().reshape(Nv * nA, D), "audio").view(Nv, nA) if fa is not None else torch.zeros(Nv, 0, device=dev)
    fb = [f.float() for f in (fv, fa) if f is not None]
    feat_b = (torch.cat(fb, dim=1) if len(fb) > 1 else fb[0]).contiguous()
    maskA = (txt_tokens != 0).to(torch.uint8).contiguous()
    name = "tva" if (fv is not None and fa is not None) else ("tv" if fv is not None else "ta")
    scores = Fn.FineSimFn.apply(feat_t.reshape(Nt * T, D), feat_b.reshape(-1, D), w_t.float(), w_v.float(), w_a.float(), maskA,
                                (Nt, Nv, T, nV, nA), [name], model.compute_dtype != torch.float32)
    return scores[0]


@torch.no_grad()
def compute_metric_ret(model, score, ids, ids_txt, dual_softmax=False, evaluate_ret_text=False):
    """test.py:714-775 on the device.  score [len(ids_txt), len(ids)]; ids / ids_txt: video id of every candidate /
    of every caption.  Returns the reference's eval_log dict."""
    assert score.shape == (len(ids_txt), len(ids))
    pos = {v: i for i, v in enumerate(ids)}
    gt = torch.tensor([pos[t] for t in ids_txt], dtype=torch.int32, device=score.device)
    score = score.float().contiguous()
    fwd = K.dual_softmax(score, model.contra_temp.data.view(1), 0) if dual_softmax else score
    rank = K.retrieval_rank(fwd, gt).float()

    def summary(r, n, prefix):
        r1, r5, r10 = [(r < k).sum().item() / n for k in (1, 5, 10)]
        return {f"{prefix}_recall": f"{round(r1 * 100, 1)}/{round(r5 * 100, 1)}/{round(r10 * 100, 1)}",
                f"{prefix}_ravg": round((r1 + r5 + r10) / 3 * 100, 1),
                f"{prefix}_medianR": torch.median(r).item() + 1, f"{prefix}_meanR": torch.mean(r).item() + 1}

    log = summary(rank, len(ids_txt), "forward")
    if evaluate_ret_text:
        bwd = K.dual_softmax(score, model.contra_temp.data.view(1), 1) if dual_softmax else score
        # rank of each caption inside its own video's column, then the best caption per video (test.py:744-752)
        colT = bwd.t().contiguous()[gt.long()].contiguous()                       # row t = column gt[t] of the score matrix
        col_rank = K.retrieval_rank(colT, torch.arange(len(ids_txt), dtype=torch.int32, device=score.device))
        best = torch.full((len(ids),), 1 << 30, dtype=torch.int32, device=score.device)
        best.scatter_reduce_(0, gt.long(), col_rank, reduce="amin")
        log.update(summary(best.float(), len(ids), "backward"))
    return log


@torch.no_grad()
def validate_ret(model, batches, task_str, evaluate_ret_text=False, dual_softmax=False):
    """test.py:249-411 for one process: run `model(batch, task, compute_loss=False)` over the loader's batches, score
    every requested modality group and return {'t_v': log, 't_va': log, 't_a': log}."""
    task = task_str.split("%")[1:]
    was_training = model.training
    model.eval()
    ft, fv, fa, toks, ids, ids_txt = [], [], [], [], [], []
    for batch in batches:
        ev = model(batch, task_str, compute_loss=False)
        ft.append(ev["feat_t"]); toks.append(ev["txt_tokens"])
        if ev["feat_v"] is not None:
            fv.append(ev["feat_v"])
        if ev["feat_a"] is not None:
            fa.append(ev["feat_a"])
        ids += list(batch["ids"])
        ids_txt += list(batch.get("ids_txt", batch["ids"]))
    model.train(was_training)
    feat_t, tokens = torch.cat(ft), torch.cat(toks)
    feat_v = torch.cat(fv) if fv else None
    feat_a = torch.cat(fa) if fa else None
    out = {}
    for group, key in (("tv", "t_v"), ("tva", "t_va"), ("ta", "t_a")):
        if group in task:
            s = _fine_scores(model, feat_t, feat_v, feat_a, tokens, group)
            out[key] = compute_metric_ret(model, s, ids, ids_txt, dual_softmax, evaluate_ret_text)
    return out
