"""VALOR pretraining model — mirror of `model/pretrain.py` (class VALOR :64-134, forward_pt
:214-541, compute_fine_matrix :178-211).  `VALOR(opts)`, `VALOR.from_pretrained(opts, sd)` and
`VALOR.forward(batch, task, compute_loss=True) -> dict` keep the reference's signatures; the state
dict keeps the reference's keys.
"""
import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .distributed import ddp_allgather, ddp_allgather_with_grads, mark
from .functional import lin_of
from .modeling import VALORModel, default_opts  # noqa: F401
from .videoswin import _Linear


class Contra_head(nn.Module):
    """pretrain.py:33-38"""

    def __init__(self, input_dim, contra_dim):
        super().__init__()
        self.linear = _Linear(input_dim, contra_dim, bias=False)


class VALOR(VALORModel):
    def __init__(self, opts):
        super().__init__(opts)
        config = opts
        self.contra_type = config.contra_type
        self.caption_type = config.caption_type
        self.contra_loss_ratio = config.contra_loss_ratio
        self.use_task_prompt = config.use_task_prompt
        self.late_fusion = config.late_fusion
        self.full_masker = config.full_masker
        if self.contra_type != "fine" or self.caption_type != "unimlm" or self.use_task_prompt or self.late_fusion \
                or self.full_masker:
            raise NotImplementedError("only the shipped pretraining recipe (fine contrastive + unimlm caption, no "
                                      "task prompt) is on the hot path")
        contra_dim = config.contra_dim
        self.contra_head_t = Contra_head(self.txt_dim, contra_dim)
        self.contra_head_v = Contra_head(self.video_dim, contra_dim)
        self.contra_head_a = Contra_head(self.audio_dim, contra_dim)
        for nm in ("text", "video", "audio"):  # pretrain.py:103-112 (index 1 is the ReLU)
            seq = nn.Sequential(_Linear(contra_dim, contra_dim), nn.Identity(), _Linear(contra_dim, 1))
            setattr(self, f"{nm}_fine_weight", seq)
        self.contra_temp = nn.Parameter(torch.tensor(0.07))
        self.contra_dim = contra_dim

    # ------------------------------------------------------------------------------------
    def forward(self, batch, task, compute_loss=True):
        if not (task.startswith("pt") or task.startswith("ret")):
            raise NotImplementedError("cap / qa decoding heads are outside the scope table (SURVEY.md §8f N3)")
        eff = task if task.startswith("pt") else "pt_contra%" + "%".join(task.split("%")[1:])
        if self.store is not None and self.training:
            cache = self.__dict__.setdefault("_unused_cache", {})
            if eff not in cache:
                cache[eff] = self.unused_parameter_names(eff)
            self.store.set_unused(cache[eff])
            if self.rng.active:
                self.rng.begin_step()     # new masks every forward (also under CUDA-graph replay: see RngState)
        if task.startswith("pt"):
            return self.forward_pt(batch, task, compute_loss=compute_loss)
        return self.forward_ret(batch, task, compute_loss=compute_loss)

    def forward_ret(self, batch, task, compute_loss=True):
        """pretrain.py:544-711: the retrieval head = the contrastive branch of forward_pt alone.  compute_loss=False
        returns {'feat_t','feat_v','feat_a','txt_tokens'} for valor_b200.retrieval.validate_ret; compute_loss=True the
        fine-grained contrastive loss over the requested groups WITHOUT the pretraining loss ratio (:699)."""
        groups = task.split("%")[1:]
        pt_task = "pt_contra%" + "%".join(groups)
        out = self.forward_pt(batch, pt_task, compute_loss=compute_loss)
        if compute_loss:
            return {"contra_loss": out["contra_loss"] / self.contra_loss_ratio}
        return out

    def unused_parameter_names(self, task):
        """Parameters that receive no gradient under `task` (the reference runs DDP with
        find_unused_parameters=True for exactly these, train_utils.py:232; its AdamW skips them,
        optim/adamw.py:52-53): the BERT pooler never runs, the prompt embedding is idle without task prompts
        (bert.py:207-213), and a modality absent from every objective leaves its tower and heads untouched."""
        used = "".join(x for t in task.split("_") for x in t.split("%")[1:])
        contra = "".join(x for t in task.split("_") if "contra" in t for x in t.split("%")[1:])
        caption = "".join(x for t in task.split("_") if "caption" in t for x in t.split("%")[1:])
        dead = ["multimodal_encoder.pooler.", "multimodal_encoder.embeddings.prompt_embedding."]
        if "v" not in used:
            dead += ["video_encoder.", "hidden_trans_video_multimodal.", "video_frame_embedding", "video_type_embeddings"]
        if "a" not in used:
            dead += ["audio_encoder.", "audio_embeddings.", "hidden_trans_audio_multimodal.", "audio_frame_embedding",
                     "audio_type_embeddings"]
        if "v" not in contra:
            dead += ["contra_head_v.", "video_fine_weight."]
        if "a" not in contra:
            dead += ["contra_head_a.", "audio_fine_weight."]
        if not contra:
            dead += ["contra_head_t.", "text_fine_weight.", "contra_temp"]
        if "v" not in caption:
            dead += ["hidden_trans_video_multimodal.", "video_frame_embedding", "video_type_embeddings"]
        if "a" not in caption:
            dead += ["audio_frame_embedding", "audio_type_embeddings"]
        if not caption:
            dead += ["cls.dense.", "cls.layernorm.", "cls.decoder.bias"]
            dead += [f"multimodal_encoder.encoder.layer.{i}.cross_attn." for i in range(len(self.multimodal_encoder.encoder.layer))]
        names = [n for n, _ in self.named_parameters()]
        return sorted(n for n in names if any(n.startswith(d) or n == d for d in dead))

    def _fine_weight(self, feat, name):
        seq = getattr(self, f"{name}_fine_weight")
        h = Fn.linear(Fn.CastFn.apply(feat, self.compute_dtype), lin_of(seq[0].weight, seq[0].bias), act=K.ACT_RELU)
        return Fn.linear(h, lin_of(seq[2].weight, seq[2].bias), out_dtype=torch.float32)  # [rows, 1] fp32

    def forward_pt(self, batch, task, compute_loss=True):
        """pretrain.py:214-541.  compute_loss=True -> {'contra_loss','caption_loss'}; compute_loss=False -> the
        reference's evaluation dict (:402-406,:446,:480,:541): L2-normalised contrastive features (local rows, no
        all-gather), the contrastive tokens, per-pass masked-token caption scores and the caption labels."""
        contra_task, caption_task = [], []
        for t in task.split("_"):
            if "mlm" in t:
                # pretrain.py:487-519 drives the mlm objective through task prompts built by the BERT tokenizer
                # (modeling.py:355-369) regardless of use_task_prompt; tokenizers and the prompt path are outside the
                # scope table, and the shipped pretraining task string has no mlm objective
                raise NotImplementedError("the mlm objective needs the tokenizer-built task prompts (out of scope)")
            elif "caption" in t:
                caption_task = t.split("%")[1:]
            elif "contra" in t:
                contra_task = t.split("%")[1:]
        txt_tokens = batch["txt_tokens"]["bert_tokens"]
        video_pixels = batch.get("video_pixels")
        audio_spectrograms = batch.get("audio_spectrograms")
        loss_dict = {}
        used = "".join(caption_task + contra_task)
        video_output = self.forward_video_encoder(video_pixels) if "v" in used else None
        audio_output = self.forward_audio_encoder(audio_spectrograms) if "a" in used else None
        # everything created from here on (fusion BERT, heads) has finished its backward when these fire: the
        # gradient all-reduce of those parameters starts there and overlaps the encoders' backward (distributed.py)
        video_output = mark(video_output, "post") if video_output is not None else None
        audio_output = mark(audio_output, "post") if audio_output is not None else None
        B, T = txt_tokens.shape
        dt = self.compute_dtype

        if contra_task:
            txt_output = self.forward_txt_encoder(txt_tokens)                                   # [B,T,768]
            # the contrastive head runs in fp32 from the projection on (features, L2 normalisation, all-gather, fine
            # similarity): the loss divides similarities by temp = 0.07, so bf16-rounded unit features alone would cost
            # ~2e-3 of the loss; the GEMMs still run on the bf16 tensor cores (split operands, FineSimFn)
            f32 = torch.float32
            feat_t = Fn.L2NormFn.apply(Fn.linear(txt_output.reshape(B * T, -1), lin_of(self.contra_head_t.linear.weight), out_dtype=f32))
            feat_t = feat_t.view(B, T, -1)
            if compute_loss:
                feat_t = ddp_allgather_with_grads.apply(feat_t)
            tokens_g = ddp_allgather(txt_tokens) if compute_loss else txt_tokens
            Na = feat_t.shape[0]
            nV = nA = 0
            feat_v = feat_a = None
            if "v" in "".join(contra_task):
                _, nV, X, C = video_output.shape
                pooled = Fn.MeanPoolFn.apply(video_output.reshape(-1, C), B * nV, X)        # modeling.py:389
                feat_v = Fn.L2NormFn.apply(Fn.linear(pooled, lin_of(self.contra_head_v.linear.weight), out_dtype=f32))
                feat_v = feat_v.view(B, nV, -1)
                if compute_loss:
                    feat_v = ddp_allgather_with_grads.apply(feat_v)
            if "a" in "".join(contra_task):
                _, nA, X, C = audio_output.shape
                cls = Fn.SelectFirstFn.apply(audio_output.reshape(-1, C), B * nA, X)        # modeling.py:399
                feat_a = Fn.L2NormFn.apply(Fn.linear(cls, lin_of(self.contra_head_a.linear.weight), out_dtype=f32))
                feat_a = feat_a.view(B, nA, -1)
                if compute_loss:
                    feat_a = ddp_allgather_with_grads.apply(feat_a)
            if not compute_loss:
                loss_dict.update(feat_t=feat_t, feat_v=feat_v, feat_a=feat_a, txt_tokens=tokens_g)      # pretrain.py:402-406
            else:
                D = feat_t.shape[-1]
                w_t = self._fine_weight(feat_t.reshape(Na * T, D), "text").view(Na, T)
                dev = feat_t.device
                w_v = self._fine_weight(feat_v.reshape(Na * nV, D), "video").view(Na, nV) if feat_v is not None else \
                    torch.zeros(Na, 0, device=dev)
                w_a = self._fine_weight(feat_a.reshape(Na * nA, D), "audio").view(Na, nA) if feat_a is not None else \
                    torch.zeros(Na, 0, device=dev)
                parts = [f for f in (feat_v, feat_a) if f is not None]
                feat_va = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]                  # pretrain.py:324
                maskA = (tokens_g != 0).to(torch.uint8).contiguous()                               # pretrain.py:304
                groups = [g for g in ("tva", "tv", "ta") if g in contra_task]                      # order of pretrain.py:397
                scores = Fn.FineSimFn.apply(feat_t.reshape(Na * T, D), feat_va.reshape(-1, D), w_t, w_v, w_a, maskA,
                                            (Na, Na, T, nV, nA), groups, dt != torch.float32)
                lo = [Fn.ContrastiveFn.apply(scores[i], self.contra_temp) for i in range(len(groups))]
                loss_dict["contra_loss"] = (sum(lo) / len(lo) * self.contra_loss_ratio).reshape(())

        if caption_task:
            media, Sv, Sa = self.media_tokens(video_output, audio_output)                       # modeling.py:485-502
            if batch.get("caption_mask") is not None:
                # the TokenMasker draw (host Python RNG in the reference, modeling.py:134-174) taken in
                # the input pipeline from the host copy of the tokens: no device->host sync in the step
                txt_input, txt_labels = batch["caption_mask"]
            else:
                txt_input, txt_labels = self.text_masker(txt_tokens, 0.6)                       # pretrain.py:428
            names = [n for n in ("tva", "tv", "ta") if n in caption_task]
            ranges = {"tva": (0, Sv + Sa), "tv": (0, Sv), "ta": (Sv, Sa)}
            npass = len(names)
            tok_all = txt_input.repeat_interleave(npass, 0)            # sample-major: row b*npass + pass
            h = self.multimodal_encoder.encode(tok_all, [True] * npass, media, [ranges[n] for n in names], B)
            logits = self.cls(h)                                                                # every position;
            # labels == -1 rows are ignored.  Every pass shares the same labels, so the mean over all
            # npass*B*T rows equals the reference's mean of per-pass means (pretrain.py:473-479).
            labels = txt_labels.repeat_interleave(npass, 0).reshape(-1)
            if getattr(self, "debug_capture", None) is not None:   # parity tests: the masked-token logits
                self.debug_capture.update(logits=logits.detach().clone(), labels=labels, npass=npass, names=names)
            if compute_loss:
                loss_dict["caption_loss"] = Fn.XentFn.apply(logits, labels).reshape(())
            else:   # pretrain.py:446,463,480,485: scores of the masked positions, per pass, plus the labels
                lg = logits.view(B, npass, T, -1)
                sel = txt_labels != -1
                for i, n in enumerate(names):
                    loss_dict[f"caption_scores_{n}"] = lg[:, i][sel]
                loss_dict["txt_labels_caption"] = txt_labels
        return loss_dict
