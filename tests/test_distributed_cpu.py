"""world_size-2 gloo tests (CPU) of the N>1 host logic: contrastive all-gather with slice-back
gradient (utils/distributed.py:38-72 semantics), flat gradient all-reduce, and a 2-rank tiny
pretraining step through the test-only CPU kernel backend."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _install_cpu_backend():
    import types
    from tests import cpu_backend
    import valor_b200.kernels as K
    for name in dir(cpu_backend):
        obj = getattr(cpu_backend, name)
        if isinstance(obj, types.FunctionType) and not name.startswith("_"):
            setattr(K, name, obj)


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from valor_b200.distributed import ddp_allgather, ddp_allgather_with_grads, allreduce_grads
        if mode == "gather":
            x = (torch.arange(6, dtype=torch.float32).view(3, 2) + 10 * rank).requires_grad_(True)
            y = ddp_allgather_with_grads.apply(x)
            w = torch.arange(1, 13, dtype=torch.float32).view(6, 2)
            (y * w).sum().backward()
            out[rank] = (y.detach().clone(), x.grad.clone(), ddp_allgather(torch.tensor([rank, rank + 5])))
        elif mode in ("step", "step_overlap"):
            _install_cpu_backend()
            from tests.test_host_logic import build
            torch.set_num_threads(2)
            cfg = dict(geom="tiny", B=2, F=2, A=1, T=16, weight_seed=0, batch_seed=123 + rank, mask_seed=1234 + rank)
            model, batch = build(cfg)
            if mode == "step_overlap":   # buckets start from inside backward (marks are placed by the forward pass)
                from valor_b200.distributed import overlap_grad_allreduce
                assert overlap_grad_allreduce(model.store) is not None
            losses = model(batch, "pt_contra%tva%tv%ta_caption%tva%tv%ta", True)
            model.store.zero_grad()
            sum(losses.values()).backward()
            local_norm = model.store.grad.norm().item()
            allreduce_grads(model.store)
            extra = model.store.reducer.last_launched_in_backward if mode == "step_overlap" else -1
            out[rank] = ({k: v.item() for k, v in losses.items()}, local_norm, model.store.grad.norm().item(),
                         model.store.grad[:1000].clone(), model.store.grad.clone(), extra)
    finally:
        dist.destroy_process_group()


def _run(mode):
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [mp.Process(target=_worker, args=(r, 2, port, mode, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return dict(out)


def test_allgather_with_grads_world2():
    out = _run("gather")
    y0, g0, t0 = out[0]
    y1, g1, t1 = out[1]
    expect = torch.cat([torch.arange(6.).view(3, 2), torch.arange(6.).view(3, 2) + 10])
    torch.testing.assert_close(y0, expect)
    torch.testing.assert_close(y1, expect)
    w = torch.arange(1, 13, dtype=torch.float32).view(6, 2)
    torch.testing.assert_close(g0, w[:3])     # backward keeps only the local rows (distributed.py:66-71)
    torch.testing.assert_close(g1, w[3:])
    assert t0.tolist() == [0, 5, 1, 6] and t1.tolist() == [0, 5, 1, 6]


@pytest.mark.slow
def test_two_rank_step_is_consistent():
    out = _run("step")
    l0, n0, a0, h0 = out[0][:4]
    l1, n1, a1, h1 = out[1][:4]
    # every rank evaluates the full global contrastive loss (identical), caption losses are local
    assert abs(l0["contra_loss"] - l1["contra_loss"]) < 1e-6
    assert abs(l0["caption_loss"] - l1["caption_loss"]) > 1e-6
    assert abs(a0 - a1) < 1e-6 * a0            # after the all-reduce both ranks hold the same gradient
    torch.testing.assert_close(h0, h1)


def test_overlapped_bucket_allreduce_matches_single_allreduce_world2():
    """The bucketed all-reduce launched from inside backward (DDP-style overlap) leaves exactly the gradients of the
    single whole-arena all-reduce, starts most buckets before backward ends, and covers the whole arena."""
    ref = _run("step")
    got = _run("step_overlap")
    for r in (0, 1):
        assert got[r][5] >= 6, f"only {got[r][5]} ranges were reduced from inside backward"
        assert torch.equal(got[r][4], ref[r][4]), (got[r][4] - ref[r][4]).abs().max()
    assert torch.equal(got[0][4], got[1][4])


def test_gradient_buckets_follow_the_model_segments():
    """Bucket boundaries derived from parameter names: one bucket per Swin stage, the 18-block stage 3 split in three
    (first / middle / last six blocks + the stage's PatchMerging), AST layers, everything after the encoders; the
    marks placed by the forward pass use the same names."""
    from valor_b200.distributed import default_segments, swin_bucket
    names = [f"video_encoder.layers.{st}.blocks.{b}.attn.qkv.weight" for st, d in enumerate((2, 2, 18, 2)) for b in range(d)]
    names += [f"video_encoder.layers.{st}.downsample.reduction.weight" for st in range(3)]
    names += ["video_encoder.patch_embed.proj.weight", "video_encoder.norm.weight", "audio_encoder.layer.3.ff_layer.linear1.weight",
              "multimodal_encoder.encoder.layer.0.output.dense.weight", "cls.dense.weight", "contra_temp"]
    segs = dict(default_segments(names))
    assert list(segs) == ["swin.0", "swin.1", "swin.2a", "swin.2b", "swin.2c", "swin.3", "ast", "post"]
    assert segs["swin.2a"] == tuple(f"video_encoder.layers.2.blocks.{b}." for b in range(6))
    assert segs["swin.2c"][-1] == "video_encoder.layers.2.downsample." and "video_encoder.layers.2.blocks.17." in segs["swin.2c"]
    assert segs["swin.0"][-1] == "video_encoder.layers.0.downsample."
    assert [swin_bucket(2, 18, b) for b in (0, 5, 6, 11, 12, 17)] == ["swin.2a", "swin.2a", "swin.2b", "swin.2b", "swin.2c", "swin.2c"]
    assert swin_bucket(1, 2, 1) == "swin.1"
    owned = [n for n in names if any(n.startswith(p) for ps in segs.values() for p in ps)]
    assert set(names) - set(owned) == {"video_encoder.patch_embed.proj.weight", "video_encoder.norm.weight", "contra_temp"}
