"""GPU parity tests, kernel by kernel, THROUGH THE C ABI (valor_b200.kernels -> libvalor_b200.so)
against the fp32 torch restatement of the same operator (tests/cpu_backend.py, itself pinned to
the reference golden vectors via tests/test_host_logic.py).

Tolerances: fp32 kernels 1e-4 relative; bf16 kernels are compared with the fp32 operator evaluated
on the SAME bf16-rounded inputs, tolerance 2e-2 of the output scale (bf16 has 8 mantissa bits).
"""
import math

import pytest
import torch

from tests import cpu_backend as R

pytestmark = pytest.mark.gpu


def K():
    import valor_b200.kernels as k
    return k


def dev(t, dtype=None):
    if t is None:
        return None
    t = t.cuda()
    return t.to(dtype) if dtype is not None and t.is_floating_point() else t


def close(got, ref, dtype, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    scale = ref.abs().max().item() + 1e-6
    tol = 2.5e-2 if dtype == torch.bfloat16 else 2e-4
    err = (got - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


def close_fro(got, ref, what="", tol=6e-3):
    """relative Frobenius error (bf16 rounding of the output alone is ~2e-3)"""
    got, ref = got.float().cpu().double(), ref.float().cpu().double()
    err = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    assert err <= tol, f"{what}: relative Frobenius error {err:.3e} (tol {tol})"


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * scale
    return x.to(dtype).float() if dtype == torch.bfloat16 else x  # value representable in `dtype`


DTYPES = [torch.float32, torch.bfloat16]

# ------------------------------------------------------------------------------------------
# GEMM: the three forms of every Linear (forward / dgrad / wgrad), tcgen05 and SIMT back ends
# ------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (256, 384, 128), (1000, 768, 768), (130, 512, 96), (4096, 128, 512), (300, 3072, 768), (257, 30522 // 32, 768),
    (64, 1, 512), (2048, 2304, 768), (128, 192, 64), (1024, 1000, 200),
]


@pytest.mark.parametrize("M,N,K_", GEMM_SHAPES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("form", ["nt", "nn", "tn"])
def test_gemm_forms(M, N, K_, dtype, form):
    k = K()
    a_k, b_k = {"nt": (True, True), "nn": (True, False), "tn": (False, False)}[form]
    a = rnd(M, K_, dtype=dtype, seed=1) if a_k else rnd(K_, M, dtype=dtype, seed=1)
    b = rnd(N, K_, dtype=dtype, seed=2) if b_k else rnd(K_, N, dtype=dtype, seed=2)
    ref = R.gemm(a, b, a_kmajor=a_k, b_kmajor=b_k)
    got = k.gemm(dev(a, dtype), dev(b, dtype), a_kmajor=a_k, b_kmajor=b_k)
    close(got, ref, dtype, f"gemm {form} {M}x{N}x{K_}")


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
@pytest.mark.parametrize("form", ["nt", "nn", "tn"])
def test_gemm_tensor_tile_widths(bn, form):
    k = K()
    M, N, K_ = 384, 520, 328
    a_k, b_k = {"nt": (True, True), "nn": (True, False), "tn": (False, False)}[form]
    dtype = torch.bfloat16
    a = rnd(M, K_, dtype=dtype, seed=3) if a_k else rnd(K_, M, dtype=dtype, seed=3)
    b = rnd(N, K_, dtype=dtype, seed=4) if b_k else rnd(K_, N, dtype=dtype, seed=4)
    ref = R.gemm(a, b, a_kmajor=a_k, b_kmajor=b_k)
    got = k.gemm(dev(a, dtype), dev(b, dtype), a_kmajor=a_k, b_kmajor=b_k, backend=k.BACKEND_TENSOR, force_bn=bn)
    close(got, ref, dtype, f"gemm tensor bn={bn} {form}")


@pytest.mark.parametrize("form", ["nt", "nn", "tn"])
@pytest.mark.parametrize("shape", [(512, 512, 256), (1000, 520, 328), (4096, 768, 192), (136, 256, 64), (50176, 512, 512)])
@pytest.mark.parametrize("pair", [1, 2])
def test_gemm_two_cta_tiles(form, shape, pair):
    """256 x 256 tiles over a CTA pair (tcgen05.mma.cta_group::2; force_bn 1256) against the single-CTA 128 x 256 tile
    (force_bn 2256) and the fp32 reference: ragged last tiles in M and N, every operand layout."""
    k = K()
    M, N, K_ = shape
    a_k, b_k = {"nt": (True, True), "nn": (True, False), "tn": (False, False)}[form]
    dtype = torch.bfloat16
    a = rnd(M, K_, dtype=dtype, seed=3) if a_k else rnd(K_, M, dtype=dtype, seed=3)
    b = rnd(N, K_, dtype=dtype, seed=4) if b_k else rnd(K_, N, dtype=dtype, seed=4)
    ref = R.gemm(a, b, a_kmajor=a_k, b_kmajor=b_k)
    got = k.gemm(dev(a, dtype), dev(b, dtype), a_kmajor=a_k, b_kmajor=b_k, backend=k.BACKEND_TENSOR, force_bn=pair * 1000 + 256)
    close(got, ref, dtype, f"gemm two-CTA={pair == 1} {form} {shape}")


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("force", [1256, 2256, 128, 192, 64])
def test_gemm_specialised_epilogues_every_tile(act, force):
    """The compile-time epilogue specialisations (bias + GELU with the pre-activation side output, residual add, the
    activation-gradient multiply on a TMA-staged tile, split-K fp32 accumulation of the weight gradient) on every tile
    shape: two-CTA 256 x 256 (1256), single-CTA 128 x 256 (2256), 128 x 128 / 192 / 64."""
    k = K()
    dtype = torch.bfloat16
    M, N, K_ = 1280, 768, 256
    a, b = rnd(M, K_, dtype=dtype, seed=5), rnd(N, K_, dtype=dtype, seed=6, scale=0.1)
    bias, res = rnd(N, seed=7), rnd(M, N, dtype=dtype, seed=8)
    kw = dict(backend=k.BACKEND_TENSOR, force_bn=force)
    ref, ref_pre = R.gemm(a, b, bias=bias, act=act, want_preact=True)
    got, got_pre = k.gemm(dev(a, dtype), dev(b, dtype), bias=dev(bias), act=act, want_preact=True, **kw)
    close(got, ref, dtype, "two-CTA act out")
    close_fro(got, ref, "act out")
    close_fro(got_pre, ref_pre, "preact")
    close(got_pre, ref_pre, dtype, "two-CTA preact")
    close(k.gemm(dev(a, dtype), dev(b, dtype), bias=dev(bias), residual=dev(res, dtype), **kw),
          R.gemm(a, b, bias=bias, residual=res), dtype, "two-CTA residual")
    aux = rnd(M, N, dtype=dtype, seed=9)
    bt = rnd(K_, N, dtype=dtype, seed=10, scale=0.1)
    got_g, ref_g = k.gemm(dev(a, dtype), dev(bt, dtype), b_kmajor=False, act=act, act_aux=dev(aux, dtype), **kw), R.gemm(a, bt, b_kmajor=False, act=act, act_aux=aux)
    close(got_g, ref_g, dtype, "act-grad")
    close_fro(got_g, ref_g, "act-grad")
    rows = 20000
    dy, x = rnd(rows, 768, dtype=dtype, seed=11, scale=0.1), rnd(rows, 512, dtype=dtype, seed=12)
    base = rnd(768, 512, seed=13)
    out = dev(base.clone())
    k.gemm(dev(dy, dtype), dev(x, dtype), a_kmajor=False, b_kmajor=False, out=out, accumulate=True, **kw)
    close(out, base + dy.t() @ x, dtype, "two-CTA wgrad split-K")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("force", [0, 1256, 2256, 128])
@pytest.mark.parametrize("with_res", [True, False])
def test_gemm_row_scale_epilogue(dtype, force, with_res):
    """DropPath inside the GEMM that ends a residual branch (videoswin.py:40-55,238,243): C = residual +
    scale[row // rows_per_group] * (A.B^T + bias), ragged last group, every tile shape and the fp32 SIMT path."""
    k = K()
    M, N, K_, rpg = 1200, 512, 256, 392
    a, b = rnd(M, K_, dtype=dtype, seed=31), rnd(N, K_, dtype=dtype, seed=32, scale=0.1)
    bias, res = rnd(N, seed=33), rnd(M, N, dtype=dtype, seed=34)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25])          # keep / keep_prob per sample; sample 1 dropped
    kw = dict(backend=k.BACKEND_TENSOR, force_bn=force) if (dtype == torch.bfloat16 and force) else {}
    if dtype == torch.float32 and force:
        pytest.skip("tile shapes belong to the bf16 tensor path")
    got = k.gemm(dev(a, dtype), dev(b, dtype), bias=dev(bias), residual=dev(res, dtype) if with_res else None,
                 row_scale=dev(scale), rows_per_group=rpg, **kw)
    ref = (a @ b.t() + bias) * scale.repeat_interleave(rpg)[:M, None] + (res if with_res else 0.0)
    close(got, ref, dtype, "row-scale epilogue")
    close_fro(got, ref, "row-scale epilogue", tol=6e-3 if dtype == torch.bfloat16 else 1e-5)
    dropped = got[rpg:2 * rpg].float().cpu()
    torch.testing.assert_close(dropped, (res[rpg:2 * rpg] if with_res else torch.zeros(rpg, N)).to(dtype).float(), rtol=0, atol=0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_epilogues(dtype, act):
    k = K()
    M, N, K_ = 512, 384, 256
    a, b = rnd(M, K_, dtype=dtype, seed=5), rnd(N, K_, dtype=dtype, seed=6, scale=0.1)
    bias = rnd(N, seed=7)
    res = rnd(M, N, dtype=dtype, seed=8)
    ref, ref_pre = R.gemm(a, b, bias=bias, act=act, residual=res, want_preact=True)
    got, got_pre = k.gemm(dev(a, dtype), dev(b, dtype), bias=dev(bias), act=act, residual=dev(res, dtype), want_preact=True)
    close(got, ref, dtype, "epilogue out")
    close(got_pre, ref_pre, dtype, "epilogue preact")
    # activation-gradient epilogue (dgrad through GELU / ReLU)
    aux = rnd(M, N, dtype=dtype, seed=9)
    ref2 = R.gemm(a, b, act=act, act_aux=aux)
    got2 = k.gemm(dev(a, dtype), dev(b, dtype), act=act, act_aux=dev(aux, dtype))
    close(got2, ref2, dtype, "epilogue act-grad")


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_wgrad_accumulate_splitk(dtype):
    """weight gradient: fp32 accumulate into an existing buffer, contraction over 20000 rows (split-K)."""
    k = K()
    Mrows, N, Kin = 20000, 384, 128
    dy, x = rnd(Mrows, N, dtype=dtype, seed=10, scale=0.1), rnd(Mrows, Kin, dtype=dtype, seed=11)
    base = rnd(N, Kin, seed=12)
    ref = base + dy.t() @ x
    out = dev(base.clone())
    k.gemm(dev(dy, dtype), dev(x, dtype), a_kmajor=False, b_kmajor=False, out=out, accumulate=True)
    close(out, ref, dtype, "wgrad split-K")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(20000, 384, 128, 0), (25088, 768, 768, 0), (3000, 3072, 768, 0), (777, 200, 72, 0),
                                   (6000, 30522, 768, 0), (4096, 512, 512, 128), (4096, 512, 512, 64)])
def test_gemm_wgrad_fused_bias_gradient(dtype, shape):
    """Linear backward: dW += dY^T X and db += column sums of dY from ONE launch (the bias gradient rides on sixteen
    extra accumulator columns fed by an all-ones operand); fp32 operands take the separate column-sum launch behind the
    same call.  Both accumulate into existing buffers."""
    k = K()
    Mrows, N, Kin, bn = shape
    dy, x = rnd(Mrows, N, dtype=dtype, seed=20, scale=0.1), rnd(Mrows, Kin, dtype=dtype, seed=21)
    base_w, base_b = rnd(N, Kin, seed=22), rnd(N, seed=23)
    out, db = dev(base_w.clone()), dev(base_b.clone())
    kw = dict(force_bn=bn, backend=k.BACKEND_TENSOR) if (bn and dtype == torch.bfloat16) else {}
    k.gemm(dev(dy, dtype), dev(x, dtype), a_kmajor=False, b_kmajor=False, out=out, accumulate=True, bias_grad=db, **kw)
    close(out, base_w + dy.t() @ x, dtype, "wgrad with fused bias gradient")
    ref_b = base_b.double() + dy.double().sum(0)
    err = (db.cpu().double() - ref_b).abs().max().item()
    assert err <= 1e-4 * (1.0 + dy.abs().double().sum(0).max().item()), f"bias gradient max abs err {err}"


def test_gemm_padded_pitch_vocab():
    """30522-wide logits with a row pitch padded to 30528 (MLM decoder, modeling.py:240-253)."""
    k = K()
    dtype = torch.bfloat16
    M, N, K_ = 200, 30522, 768
    a, b = rnd(M, K_, dtype=dtype, seed=13), rnd(N, K_, dtype=dtype, seed=14, scale=0.05)
    bias = rnd(N, seed=15)
    out = torch.empty(M, 30528, device="cuda", dtype=dtype)[:, :N]
    k.gemm(dev(a, dtype), dev(b, dtype), bias=dev(bias), out=out)
    close(out, R.gemm(a, b, bias=bias), dtype, "vocab logits")
    # dgrad with K = 30522 and wgrad with M' = 30522 read the padded-pitch matrix
    g = out  # any matrix with that layout
    close(k.gemm(g, dev(b, dtype), b_kmajor=False), R.gemm(g.float().cpu(), b, b_kmajor=False), dtype, "vocab dgrad")
    wg = torch.zeros(N, K_, device="cuda")
    k.gemm(g, dev(a, dtype), a_kmajor=False, b_kmajor=False, out=wg, accumulate=True)
    close(wg, g.float().cpu().t() @ a, dtype, "vocab wgrad")


# ------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N", [(1000, 128), (333, 768), (64, 1024), (50, 2048), (77, 512), (10, 260)])
def test_layernorm(dtype, M, N):
    k = K()
    x, dy = rnd(M, N, dtype=dtype, seed=1, scale=2.0) + 0.5, rnd(M, N, dtype=dtype, seed=2)
    x = x.to(dtype).float()
    g, b = rnd(N, seed=3) * 0.3 + 1.0, rnd(N, seed=4) * 0.1
    for eps in (1e-5, 1e-12):
        y_r, mean_r, rstd_r = R.layernorm_fwd(x, g, b, eps)
        y, mean, rstd = k.layernorm_fwd(dev(x, dtype), dev(g), dev(b), eps)
        close(y, y_r, dtype, "ln fwd")
        close(mean, mean_r, torch.float32, "ln mean")
        dg_r, db_r = torch.zeros(N), torch.zeros(N)
        dx_r = R.layernorm_bwd(dy, x, g, mean_r, rstd_r, dg_r, db_r)
        dg, db = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
        dx = k.layernorm_bwd(dev(dy, dtype), dev(x, dtype), dev(g), mean, rstd, dg, db)
        dres = rnd(M, N, dtype=dtype, seed=9)
        dx2 = k.layernorm_bwd(dev(dy, dtype), dev(x, dtype), dev(g), mean, rstd, torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), dres=dev(dres, dtype))
        close(dx2, dx_r.float() + dres, dtype, "ln dx + residual grad")
        close(dx, dx_r, dtype, "ln dx")
        close(dg, dg_r, dtype, "ln dgamma")
        close(db, db_r, dtype, "ln dbeta")


@pytest.mark.parametrize("dtype", DTYPES)
def test_l2norm(dtype):
    k = K()
    x, dy = rnd(300, 512, dtype=dtype, seed=1), rnd(300, 512, dtype=dtype, seed=2)
    y_r, n_r = R.l2norm_fwd(x)
    y, n = k.l2norm_fwd(dev(x, dtype))
    close(y, y_r, dtype, "l2 fwd")
    close(k.l2norm_bwd(dev(dy, dtype), dev(x, dtype), n), R.l2norm_bwd(dy, x, n_r), dtype, "l2 bwd")


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", ["bert_self", "ast", "cross", "cross_merged"])
def test_mha(dtype, case):
    k = K()
    H, hd = 12, 64
    Hd = H * hd
    if case == "bert_self":
        P_, Nq, nk = 6, 32, 32
        qkv = rnd(P_ * Nq, 3 * Hd, dtype=dtype, seed=1)
        q, kk, v = qkv[:, :Hd], qkv[:, Hd:2 * Hd], qkv[:, 2 * Hd:]
        lens = torch.tensor([32, 20, 9, 32, 15, 3])
        key_valid = (torch.arange(nk)[None] < lens[:, None]).to(torch.uint8)
        causal = torch.tensor([0, 0, 0, 1, 1, 1], dtype=torch.uint8)
        kw = dict(key_valid=key_valid, causal=causal)
    elif case == "ast":
        P_, Nq, nk = 3, 129, 129
        qkv = rnd(P_ * Nq, 3 * Hd, dtype=dtype, seed=1)
        q, kk, v = qkv[:, :Hd], qkv[:, Hd:2 * Hd], qkv[:, 2 * Hd:]
        kw = {}
    elif case == "cross_merged":
        # the three caption passes of a sample as ONE problem: 3*32 queries over the sample's 650 media tokens, every
        # query restricted to its pass's key subset (tva: all, tv: video rows, ta: audio rows)
        B, S, T = 2, 650, 32
        P_, Nq, nk = B, 3 * T, S
        q = rnd(P_ * Nq, Hd, dtype=dtype, seed=1)
        kvb = rnd(B * S, 2 * Hd, dtype=dtype, seed=2)
        kk, v = kvb[:, :Hd], kvb[:, Hd:]
        ranges = [(0, 650), (0, 392), (392, 258)]
        kw = dict(q_key_range=torch.tensor([[ranges[i // T][0], ranges[i // T][0] + ranges[i // T][1]] for i in range(Nq)],
                                           dtype=torch.int32))
    else:
        B, S, Nq = 2, 100, 32
        P_ = 3 * B
        q = rnd(P_ * Nq, Hd, dtype=dtype, seed=1)
        kvb = rnd(B * S, 2 * Hd, dtype=dtype, seed=2)
        kk, v = kvb[:, :Hd], kvb[:, Hd:]
        ranges = [(0, 100), (0, 60), (60, 40)]
        kw = dict(kv_row0=torch.tensor([b * S + st for st, ln in ranges for b in range(B)], dtype=torch.int32),
                  kv_len=torch.tensor([ln for st, ln in ranges for b in range(B)], dtype=torch.int32))
        nk = 100
    scale = 1 / math.sqrt(hd)
    do = rnd(P_ * Nq, Hd, dtype=dtype, seed=3)
    o_r, lse_r = R.mha_fwd(q, kk, v, P_, H, hd, Nq, nk, scale, **kw)
    dq_r = torch.zeros(P_ * Nq, Hd)
    dkv_r = R.mha_bwd(q, kk, v, o_r, do, lse_r, dq_r, P_, H, hd, Nq, nk, scale, **kw)
    kwd = {a: dev(b) for a, b in kw.items()}
    if case in ("cross", "cross_merged"):
        qd, kvd = dev(q, dtype), dev(kvb, dtype)
        kd, vd = kvd[:, :Hd], kvd[:, Hd:]
    else:
        qkvd = dev(qkv, dtype)
        qd, kd, vd = qkvd[:, :Hd], qkvd[:, Hd:2 * Hd], qkvd[:, 2 * Hd:]
    o, lse = k.mha_fwd(qd, kd, vd, P_, H, hd, Nq, nk, scale, **kwd)
    close(o, o_r, dtype, "mha o")
    close(lse, lse_r, torch.float32 if dtype == torch.float32 else dtype, "mha lse")
    dq = torch.empty(P_ * Nq, Hd, device="cuda", dtype=dtype)
    dkv = k.mha_bwd(qd, kd, vd, o, dev(do, dtype), lse, dq, P_, H, hd, Nq, nk, scale, **kwd)
    close(dq, dq_r, dtype, "mha dq")
    close(dkv, dkv_r, dtype, "mha dkv")
    if case != "cross":  # every K/V row has one owner -> direct outputs in the compute dtype
        dkv2 = torch.full((kd.shape[0], 2 * Hd), 7.0, device="cuda", dtype=dtype)
        k.mha_bwd(qd, kd, vd, o, dev(do, dtype), lse, dq, P_, H, hd, Nq, nk, scale, dkv_out=(dkv2[:, :Hd], dkv2[:, Hd:]), **kwd)
        close(dkv2, dkv_r, dtype, "mha dkv direct")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("grid,win,shift,heads", [
    ((2, 2, 14, 14), (2, 7, 7), (0, 0, 0), 4),
    ((2, 2, 14, 14), (2, 7, 7), (0, 3, 3), 4),
    ((1, 8, 14, 7), (8, 7, 7), (0, 3, 0), 2),
    ((1, 16, 7, 7), (8, 7, 7), (4, 0, 0), 2),
    ((3, 4, 7, 7), (4, 7, 7), (0, 0, 0), 8),
    ((1, 16, 7, 7), (8, 7, 7), (4, 0, 0), 4),     # temporal shift: mask regions along d, one-CTA-per-window kernels
    ((2, 4, 14, 14), (4, 7, 7), (0, 3, 3), 4),    # 196-token windows, masked border windows
    ((1, 8, 21, 14), (8, 7, 7), (0, 3, 3), 1),    # 392-token windows, 3x2 window grid, single head
])
@pytest.mark.parametrize("backend", ["auto", "tensor", "mma_sync"])
def test_window_attention(dtype, grid, win, shift, heads, backend):
    """shift / partition / relative-position bias / -100 mask evaluated in-kernel vs the reference's
    roll + window_partition + bias gather + compute_mask (videoswin.py:75-84,137-163,272-285).
    heads == 2 cases run head dim 64 (the key-blocked flash kernels), the others head dim 32
    (the one-CTA-per-window kernels)."""
    k = K()
    if backend != "auto" and dtype != torch.bfloat16:
        pytest.skip("backend selection only matters for the bf16 tensor-core paths")
    # tensor: tcgen05 forward + backward; auto: mma.sync forward + tcgen05 backward (the faster pair); mma_sync: round-1 kernels
    be = {"auto": k.BACKEND_AUTO, "tensor": k.BACKEND_TENSOR, "mma_sync": k.BACKEND_MMA_SYNC}[backend]
    hd_ = 64 if (heads == 2 and grid[1] == 16) else 32
    if backend == "tensor" and hd_ != 32:
        pytest.skip("the tcgen05 window kernels are head-dim-32 kernels")
    hd = 64 if (heads == 2 and grid[1] == 16) else 32
    C = heads * hd
    cfg_win = (8, 7, 7)
    tokens = grid[0] * grid[1] * grid[2] * grid[3]
    qkv = rnd(tokens, 3 * C, dtype=dtype, seed=1)
    table = rnd((2 * 8 - 1) * 13 * 13, heads, seed=2, scale=0.5)
    do = rnd(tokens, C, dtype=dtype, seed=3)
    scale = hd ** -0.5
    o_r, lse_r = R.window_attn_fwd(qkv, table, grid, win, shift, cfg_win, heads, hd, scale)
    dt_r = torch.zeros_like(table)
    dqkv_r = R.window_attn_bwd(qkv, o_r, do, lse_r, table, dt_r, grid, win, shift, cfg_win, heads, hd, scale)
    o, lse = k.window_attn_fwd(dev(qkv, dtype), dev(table), grid, win, shift, cfg_win, heads, hd, scale, backend=be)
    close(o, o_r, dtype, "window o")
    close(lse.reshape(-1), lse_r.reshape(-1), torch.float32 if dtype == torch.float32 else dtype, "window lse")
    dt = torch.zeros_like(table).cuda()
    dqkv = k.window_attn_bwd(dev(qkv, dtype), o, dev(do, dtype), lse, dev(table), dt, grid, win, shift, cfg_win, heads,
                             hd, scale, backend=be)
    close(dqkv, dqkv_r, dtype, "window dqkv")
    close(dt, dt_r, dtype, "window dtable")


# ------------------------------------------------------------------------------------------
# data movement
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_im2col_and_assemble(dtype):
    k = K()
    video = rnd(2, 3, 3, 32, 48, seed=1)
    close(k.swin_im2col(dev(video), dtype), R.swin_im2col(video, torch.float32), dtype, "swin im2col")
    spec = rnd(3, 64, 96, seed=2)
    close(k.audio_im2col(dev(spec), 16, dtype), R.audio_im2col(spec, 16, torch.float32), dtype, "audio im2col")
    BA, Pn, Hd = 3, 24, 768
    tok, cls, pos = rnd(BA * Pn, Hd, dtype=dtype, seed=3), rnd(Hd, seed=4), rnd(Pn + 1, Hd, seed=5)
    close(k.ast_assemble_fwd(dev(tok, dtype), dev(cls), dev(pos), BA, Pn), R.ast_assemble_fwd(tok, cls, pos, BA, Pn), dtype, "assemble")
    dx = rnd(BA * (Pn + 1), Hd, dtype=dtype, seed=6)
    dc_r, dp_r = torch.zeros(Hd), torch.zeros(Pn + 1, Hd)
    dt_r = R.ast_assemble_bwd(dx, dc_r, dp_r, BA, Pn)
    dc, dp = torch.zeros(Hd, device="cuda"), torch.zeros(Pn + 1, Hd, device="cuda")
    close(k.ast_assemble_bwd(dev(dx, dtype), dc, dp, BA, Pn), dt_r, dtype, "assemble dtok")
    close(dc, dc_r, dtype, "assemble dcls")
    close(dp, dp_r, dtype, "assemble dpos")


@pytest.mark.parametrize("dtype", DTYPES)
def test_embed_media_merge_pool(dtype):
    k = K()
    V, Hd, Rr, Tn = 1000, 768, 5, 16
    tokens = torch.randint(0, V, (Rr, Tn), generator=torch.Generator().manual_seed(1))
    word, pos, typ = rnd(V, Hd, seed=2), rnd(64, Hd, seed=3), rnd(2, Hd, seed=4)
    close(k.bert_embed_fwd(dev(tokens), dev(word), dev(pos), dev(typ)[0], dtype), R.bert_embed_fwd(tokens, word, pos, typ[0], torch.float32), dtype, "embed")
    de = rnd(Rr * Tn, Hd, dtype=dtype, seed=5)
    dw_r, dp_r, dt_r = torch.zeros(V, Hd), torch.zeros(64, Hd), torch.zeros(Hd)
    R.bert_embed_bwd(de, tokens, dw_r, dp_r, dt_r)
    dw, dp, dt = torch.zeros(V, Hd, device="cuda"), torch.zeros(64, Hd, device="cuda"), torch.zeros(Hd, device="cuda")
    k.bert_embed_bwd(dev(de, dtype), dev(tokens), dw, dp, dt)
    close(dw, dw_r, dtype, "embed dword"); close(dp, dp_r, dtype, "embed dpos"); close(dt, dt_r, dtype, "embed dtype")
    # media input
    B, nf, X, S, row0 = 2, 3, 5, 25, 10
    x, fe, te = rnd(B * nf * X, Hd, dtype=dtype, seed=6), rnd(32, Hd, seed=7), rnd(Hd, seed=8)
    out_r = torch.zeros(B * S, Hd)
    R.media_input_fwd(x, fe, te, out_r, B, nf, X, S, row0)
    out = torch.zeros(B * S, Hd, device="cuda", dtype=dtype)
    k.media_input_fwd(dev(x, dtype), dev(fe), dev(te), out, B, nf, X, S, row0)
    close(out, out_r, dtype, "media fwd")
    dout = rnd(B * S, Hd, dtype=dtype, seed=9)
    df_r, dty_r = torch.zeros(32, Hd), torch.zeros(Hd)
    din_r = R.media_input_bwd(dout, df_r, dty_r, B, nf, X, S, row0)
    df, dty = torch.zeros(32, Hd, device="cuda"), torch.zeros(Hd, device="cuda")
    close(k.media_input_bwd(dev(dout, dtype), df, dty, B, nf, X, S, row0), din_r, dtype, "media din")
    close(df, df_r, dtype, "media dframe"); close(dty, dty_r, dtype, "media dtype")
    # patch merge both directions
    BD, H, W, C = 3, 6, 4, 64
    xm = rnd(BD * H * W, C, dtype=dtype, seed=10)
    ym = k.patch_merge(dev(xm, dtype), BD, H, W, C, False)
    close(ym, R.patch_merge(xm, BD, H, W, C, False), dtype, "merge")
    close(k.patch_merge(ym, BD, H, W, C, True), xm, dtype, "merge inverse round trip")
    # pooling
    xp = rnd(4 * 49, 256, dtype=dtype, seed=11)
    close(k.mean_pool_fwd(dev(xp, dtype), 4, 49), R.mean_pool_fwd(xp, 4, 49), dtype, "pool")
    dyp = rnd(4, 256, dtype=dtype, seed=12)
    close(k.mean_pool_bwd(dev(dyp, dtype), 4, 49), R.mean_pool_bwd(dyp, 4, 49), dtype, "pool bwd")
    # colsum + act_bwd
    dyc = rnd(5000, 384, dtype=dtype, seed=13)
    db = torch.zeros(384, device="cuda")
    k.colsum(dev(dyc, dtype), db)
    close(db, dyc.sum(0), dtype, "colsum")
    for act in (1, 2, 3):
        h = rnd(300, 128, dtype=dtype, seed=14)
        close(k.act_bwd(dev(dyc[:300, :128].contiguous(), dtype), dev(h, dtype), act), R.act_bwd(dyc[:300, :128], h, act), dtype, "act_bwd")


# ------------------------------------------------------------------------------------------
# losses + optimizer
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_xent(dtype):
    k = K()
    M, V = 96, 30522
    logits = rnd(M, V, dtype=dtype, seed=1, scale=2.0)
    labels = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(2))
    labels[::3] = -1
    loss_r, lse_r, acc_r = R.xent_fwd(logits, labels)
    buf = torch.empty(M, 30528, device="cuda", dtype=dtype)[:, :V]
    buf.copy_(logits)
    loss, lse, acc = k.xent_fwd(buf, dev(labels))
    assert abs(loss.item() - loss_r.item()) <= 2e-3 * abs(loss_r.item())
    g = torch.tensor([0.7])
    d_r = R.xent_bwd(logits.clone(), labels, lse_r, acc_r, g)
    d = k.xent_bwd(buf, dev(labels), lse, acc, dev(g))
    close(d, d_r, dtype, "xent dlogits")


def test_fine_similarity_and_contrastive():
    k = K()
    Na, Nb, T, nV, nA, D = 6, 6, 16, 4, 2, 64
    Vt = nV + nA
    L = rnd(Na * T, Nb * Vt, seed=1)
    lens = torch.tensor([16, 9, 4, 12, 16, 7])
    mA = (torch.arange(T)[None] < lens[:, None]).to(torch.uint8)
    wA, wB = rnd(Na, T, seed=2), rnd(Nb, Vt, seed=3)
    wsA_r = R.masked_softmax_fwd(wA, mA)
    close(k.masked_softmax_fwd(dev(wA), dev(mA)), wsA_r, torch.float32, "masked softmax")
    for v0, nv in ((0, Vt), (0, nV), (nV, nA)):
        wsB_r = R.masked_softmax_fwd(wB[:, v0:v0 + nv].contiguous(), None)
        sc_r, av_r, at_r = R.fine_reduce_fwd(L, mA, wsA_r, wsB_r, Na, Nb, T, Vt, v0, nv)
        sc, av, at = k.fine_reduce_fwd(dev(L), dev(mA), dev(wsA_r), dev(wsB_r), Na, Nb, T, Vt, v0, nv)
        close(sc, sc_r, torch.float32, "fine score")
        ds = rnd(Na, Nb, seed=4)
        dL_r, dA_r, dB_r = torch.zeros_like(L), torch.zeros_like(wsA_r), torch.zeros_like(wsB_r)
        R.fine_reduce_bwd(L, mA, wsA_r, wsB_r, ds, av_r, at_r, dL_r, dA_r, dB_r, Na, Nb, T, Vt, v0, nv)
        dL, dA, dB = torch.zeros_like(L).cuda(), torch.zeros_like(wsA_r).cuda(), torch.zeros_like(wsB_r).cuda()
        k.fine_reduce_bwd(dev(L), dev(mA), dev(wsA_r), dev(wsB_r), dev(ds), av, at, dL, dA, dB, Na, Nb, T, Vt, v0, nv)
        close(dL, dL_r, torch.float32, "fine dL"); close(dA, dA_r, torch.float32, "fine dwsA"); close(dB, dB_r, torch.float32, "fine dwsB")
    S = rnd(40, 40, seed=5, scale=0.3)
    temp = torch.tensor([0.07])
    loss_r, rl_r, cl_r = R.contrastive_fwd(S, temp)
    loss, rl, cl = k.contrastive_fwd(dev(S), dev(temp))
    assert abs(loss.item() - loss_r.item()) <= 1e-4 * abs(loss_r.item())
    g = torch.tensor([1.3])
    dt_r = torch.zeros(1)
    dS_r = R.contrastive_bwd(S, temp, rl_r, cl_r, g, dt_r)
    dt = torch.zeros(1, device="cuda")
    close(k.contrastive_bwd(dev(S), dev(temp), rl, cl, dev(g), dt), dS_r, torch.float32, "contrastive dS")
    assert abs(dt.item() - dt_r.item()) <= 1e-3 * abs(dt_r.item())


def test_adamw_against_reference_formula():
    """optim/adamw.py:50-101 restated in oracle.valor_oracle.adamw_step; clip per train_utils.py:359."""
    from oracle import valor_oracle as vo
    k = K()
    n = 100003
    p, g = rnd(n, seed=1), rnd(n, seed=2) * 3
    m, v = torch.zeros(n), torch.zeros(n)
    pr, gr, mr, vr = p.clone(), g.clone(), m.clone(), v.clone()
    total = vo.clip_grad_norm_([gr], 5.0)
    pd, gd, md, vd = dev(p), dev(g), dev(m), dev(v)
    lp = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    sumsq, norm = torch.zeros(1, device="cuda"), torch.zeros(2, device="cuda")
    for step in (1, 2, 3):
        lr = 1e-4 * step / 3
        vo.adamw_step(pr, gr, mr, vr, step, lr, weight_decay=0.01)
        hyper = torch.tensor([lr, 0.9, 0.98, 1e-6, 0.01, lr * math.sqrt(1 - 0.98 ** step) / (1 - 0.9 ** step), 0, 0]).cuda()
        sumsq.zero_()
        k.grad_sumsq(gd, sumsq)
        k.clip_coef(sumsq, 5.0, norm)
        k.adamw(pd, gd, md, vd, lp, hyper, norm[1:2])
    assert abs(norm[0].item() - total.item()) <= 1e-4 * total.item()
    close(pd, pr, torch.float32, "adamw p")
    close(lp, pr, torch.bfloat16, "adamw bf16 copy")


# ------------------------------------------------------------------------------------------
# stochastic regularisation (Dropout / DropPath): counter-based masks regenerated in the backward
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_dropout_masks_are_regenerated_and_unbiased(dtype):
    k = K()
    R, C, p = 4096, 768, 0.1
    st = torch.tensor([1234, 1 << 24], dtype=torch.int64, device="cuda")
    x = torch.ones(R, C, device="cuda", dtype=dtype)
    res = torch.full((R, C), 2.0, device="cuda", dtype=dtype)
    y = k.dropout(x, None, p, st, 7)
    keep = (y != 0).float()
    assert abs(keep.mean().item() - (1 - p)) < 3e-3                         # 3.1M draws: sigma ~ 1.7e-4
    vals = y[y != 0].float().unique()
    assert len(vals) == 1 and abs(vals.item() - 1 / (1 - p)) < 1e-2         # kept entries scaled by 1/(1-p)
    assert torch.equal(k.dropout(x, None, p, st, 7), y)                     # same {state, site} -> same mask (backward)
    y2 = k.dropout(x, res, p, st, 7)
    torch.testing.assert_close(y2.float(), y.float() + 2.0, rtol=1e-2, atol=1e-2)
    assert not torch.equal(k.dropout(x, None, p, st, 8), y)                 # another call site
    st2 = torch.tensor([1234, 2 << 24], dtype=torch.int64, device="cuda")
    assert not torch.equal(k.dropout(x, None, p, st2, 7), y)                # another step
    # rows / columns are not correlated: per-row and per-column keep rates stay near 1-p
    assert (keep.mean(dim=1) - (1 - p)).abs().max().item() < 0.06 and (keep.mean(dim=0) - (1 - p)).abs().max().item() < 0.03


def test_droppath_scale_and_row_scale():
    k = K()
    st = torch.tensor([5, 3 << 24], dtype=torch.int64, device="cuda")
    sc = k.droppath_scale(4096, 0.2, st, 3)
    assert set(torch.round(sc * 1000).long().unique().tolist()) <= {0, 1250}        # {0, 1/keep_prob}, videoswin.py:45-50
    assert abs((sc > 0).float().mean().item() - 0.8) < 0.03
    x = rnd(8 * 49, 128, dtype=torch.bfloat16, seed=1)
    r = rnd(8 * 49, 128, dtype=torch.bfloat16, seed=2)
    s8 = torch.tensor([0, 1.25, 1.25, 0, 1.25, 1.25, 1.25, 0.0], device="cuda")
    got = k.row_scale(dev(x, torch.bfloat16), s8, 49, dev(r, torch.bfloat16))
    want = x * s8.cpu().repeat_interleave(49)[:, None] + r
    close(got, want, torch.bfloat16, "row_scale")
    close(k.row_scale(dev(x, torch.bfloat16), s8, 49), x * s8.cpu().repeat_interleave(49)[:, None], torch.bfloat16, "row_scale bwd")


# ------------------------------------------------------------------------------------------
# retrieval evaluation (test.py:680-775)
# ------------------------------------------------------------------------------------------
def test_retrieval_rank_and_dual_softmax():
    k = K()
    g = torch.Generator().manual_seed(3)
    S = torch.randn(700, 333, generator=g)
    gt = torch.randint(0, 333, (700,), generator=g).int()
    close_i = lambda a, b: torch.equal(a.cpu().long(), b.long())
    assert close_i(k.retrieval_rank(S.cuda(), gt.cuda()), (S > S.gather(1, gt.long()[:, None])).sum(1))
    gc = torch.randint(0, 700, (333,), generator=g).int()
    assert close_i(k.retrieval_rank(S.cuda(), gc.cuda(), by_column=True), (S.t() > S.t().gather(1, gc.long()[:, None])).sum(1))
    temp = torch.tensor([0.07])
    for dim in (0, 1):
        want = S * torch.softmax(S / temp, dim=dim) * S.shape[dim]
        torch.testing.assert_close(k.dual_softmax(S.cuda(), temp.cuda(), dim).cpu(), want, rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_retrieval_scoring_at_the_msrvtt_size(dtype):
    """BASELINE configs[4]: 1 000 clips x 1 000 captions, text 32 tokens, 8 frame + 2 audio slots, d = 512.  The score
    matrix of the one-GEMM + max/max reduction path against the reference's einsum formulation (pretrain.py:200-209) on
    a 64 x 1000 slice in fp64, and the device ranks against sort + index."""
    import types
    from valor_b200 import retrieval as R
    from valor_b200 import functional as Fn
    k = K()
    g = torch.Generator().manual_seed(5)
    Nt = Nv = 1000
    T, nV, nA, D = 32, 8, 2, 512
    ft = torch.nn.functional.normalize(torch.randn(Nt, T, D, generator=g), dim=-1)
    fv = torch.nn.functional.normalize(torch.randn(Nv, nV, D, generator=g), dim=-1)
    fa = torch.nn.functional.normalize(torch.randn(Nv, nA, D, generator=g), dim=-1)
    lens = torch.randint(8, 31, (Nt,), generator=g)
    tokens = (torch.arange(T)[None, :] < lens[:, None]).long() * 1000
    w_t, w_v, w_a = torch.randn(Nt, T, generator=g), torch.randn(Nv, nV, generator=g), torch.randn(Nv, nA, generator=g)
    maskA = (tokens != 0).to(torch.uint8)
    split = dtype == torch.bfloat16
    sc = Fn.FineSimFn.apply(ft.reshape(-1, D).cuda(), torch.cat((fv, fa), 1).reshape(-1, D).cuda(), w_t.cuda(), w_v.cuda(), w_a.cuda(),
                            maskA.cuda(), (Nt, Nv, T, nV, nA), ["tva"], split)[0]
    torch.cuda.synchronize()
    # reference formulation on a slice (pretrain.py:193-209), float64
    A, B = ft[:64].double(), torch.cat((fv, fa), 1).double()
    wa = w_t[:64].double().masked_fill(maskA[:64] == 0, float("-inf")).softmax(-1)
    wb = torch.cat((w_v, w_a), 1).double().softmax(-1)
    L = torch.einsum("atd,bvd->abtv", A, B) * maskA[:64].double()[:, None, :, None]
    want = 0.5 * ((L.max(-1)[0] * wa[:, None, :]).sum(-1) + (L.max(-2)[0] * wb[None, :, :]).sum(-1))
    torch.testing.assert_close(sc[:64].cpu().double(), want, rtol=2e-4, atol=2e-5)
    ids = [f"v{i}" for i in range(Nv)]
    perm = torch.randperm(Nt, generator=g).tolist()
    ids_txt = [f"v{p}" for p in perm]
    model = types.SimpleNamespace(contra_temp=types.SimpleNamespace(data=torch.tensor(0.07).cuda()))
    log = R.compute_metric_ret(model, sc, ids, ids_txt, evaluate_ret_text=True)
    order = sc.cpu().sort(dim=-1, descending=True)[1]
    rank = torch.tensor([(order[i] == perm[i]).nonzero().item() for i in range(Nt)]).float()
    assert abs(log["forward_meanR"] - (rank.mean().item() + 1)) < 1e-2 and log["forward_medianR"] == rank.median().item() + 1
    r1 = (rank < 1).float().mean().item()
    assert log["forward_recall"].split("/")[0] == str(round(r1 * 100, 1))


@pytest.mark.parametrize("case", ["self", "cross"])
def test_attention_probability_dropout(case):
    """bert.py:283,334 / transformer.py:128: dropout on the softmax output inside the attention kernels.  With V = I the
    output IS the dropped probability matrix: kept entries are P/(1-p), the keep rate is 1-p, the log-sum-exp is that of
    the undropped softmax; the backward must regenerate the same mask (gradients vs the reference formulas fed with the
    extracted mask)."""
    import tests.cpu_backend as RB
    k = K()
    H, hd, pd = 2, 64, 0.1
    if case == "self":
        P_, Nq, nk = 6, 48, 48
    else:
        P_, Nq, nk = 3, 40, 64
    Hd = H * hd
    q = rnd(P_ * Nq, Hd, dtype=torch.bfloat16, seed=1)
    kk = rnd(P_ * nk, Hd, dtype=torch.bfloat16, seed=2)
    eye = torch.zeros(P_ * nk, Hd)
    for p in range(P_):
        for h in range(H):
            eye[p * nk:(p + 1) * nk, h * hd:h * hd + nk] = torch.eye(nk)
    st = torch.tensor([77, 5 << 24], dtype=torch.int64, device="cuda")
    drop = (pd, st, 9)
    scale = hd ** -0.5
    kw = dict(P_=P_, H=H, hd=hd, Nq=Nq, max_nk=nk, scale=scale)
    qd, kd, vd = dev(q, torch.bfloat16), dev(kk, torch.bfloat16), dev(eye, torch.bfloat16)
    o_drop, lse = k.mha_fwd(qd, kd, vd, drop=drop, **kw)
    o_ref, lse_ref = k.mha_fwd(qd, kd, vd, **kw)
    torch.testing.assert_close(lse, lse_ref)                                     # statistics of the undropped softmax
    assert torch.equal(k.mha_fwd(qd, kd, vd, drop=drop, **kw)[0], o_drop)         # deterministic in {state, site}
    keep = {}
    rate = []
    ncol = min(nk, hd)
    for p in range(P_):
        for h in range(H):
            pd_ = o_drop[p * Nq:(p + 1) * Nq, h * hd:h * hd + ncol].float().cpu()
            pr_ = o_ref[p * Nq:(p + 1) * Nq, h * hd:h * hd + ncol].float().cpu()
            big = pr_ > 1e-3
            m = (pd_ > 0.5 * pr_)
            torch.testing.assert_close(pd_[m & big], (pr_ / (1 - pd))[m & big], rtol=2e-2, atol=1e-3)
            assert (pd_[~m & big].abs() < 1e-6).all()
            rate.append(m[big].float().mean().item())
            full = torch.ones(Nq, nk, dtype=torch.bool)
            full[:, :ncol] = m | ~big                       # tiny probabilities: treat as kept (their gradient weight is ~0)
            keep[(p, h)] = full
    assert abs(sum(rate) / len(rate) - (1 - pd)) < 0.02
    if nk > hd:
        return      # the mask of the key columns beyond the 64 identity columns cannot be read back: forward checks only
    # backward with the extracted mask
    v = rnd(P_ * nk, Hd, dtype=torch.bfloat16, seed=3)
    do = rnd(P_ * Nq, Hd, dtype=torch.bfloat16, seed=4)
    vd2 = dev(v, torch.bfloat16)
    o2, lse2 = k.mha_fwd(qd, kd, vd2, drop=drop, **kw)
    dq = torch.empty(P_ * Nq, Hd, device="cuda", dtype=torch.bfloat16)
    dkv = torch.empty(P_ * nk, 2 * Hd, device="cuda", dtype=torch.bfloat16)
    k.mha_bwd(qd, kd, vd2, o2, dev(do, torch.bfloat16), lse2, dq, P_, H, hd, Nq, nk, scale, dkv_out=(dkv[:, :Hd], dkv[:, Hd:]), drop=drop)
    RB.KEEP_OVERRIDE = keep
    try:
        o_r, lse_r = RB.mha_fwd(q, kk, v, P_, H, hd, Nq, nk, scale, drop=(pd, None, 0))
        dq_r = torch.zeros(P_ * Nq, Hd)
        dkv_r = RB.mha_bwd(q, kk, v, o_r, do, lse_r, dq_r, P_, H, hd, Nq, nk, scale, drop=(pd, None, 0))
    finally:
        RB.KEEP_OVERRIDE = None
    close(o2, o_r, torch.bfloat16, "dropout o")
    close(dq, dq_r, torch.bfloat16, "dropout dq")
    close(dkv, dkv_r, torch.bfloat16, "dropout dkv")
