"""LR schedule and step driver — mirror of `optim/sched.py:27-41` and the step tail of
`train_utils.conduct_train` (:344-363).  The AdamW arithmetic itself is the fused arena kernel
(params.ParamStore.optimizer_step -> valor_adamw)."""


def warmup_linear(x, warmup_ratio):
    """optim/sched.py:27-32"""
    if x < warmup_ratio:
        return x / warmup_ratio
    return max((x - 1.0) / (warmup_ratio - 1.0), 0)


def get_lr_sched(global_step, opts):
    """optim/sched.py:37-41"""
    return warmup_linear(global_step / opts.num_train_steps, opts.warmup_ratio)
