"""VideoSwin encoder — host-side mirror of the reference's `model/videoswin.py`.

Same class names, constructor arguments and state-dict keys as the reference
(`SwinTransformer3D` videoswin.py:378-439; `BasicLayer` :287-345; `SwinTransformerBlock3D`
:165-245; `WindowAttention3D` :101-163; `PatchMerging` :247-270; `PatchEmbed3D` :347-376),
but the data path is B200-native: activations live as one [tokens, C] matrix in the natural
(b, d, h, w) order for the whole tower; torch.roll / window_partition / window_reverse and the
bias/mask tensors never materialise — the window-attention kernel indexes them on the fly;
LayerNorm, GEMM(+bias+GELU / +residual) are the only other launches of a block.
"""
import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from .functional import LN, lin_of
from .distributed import mark, swin_bucket


def get_window_size(x_size, window_size, shift_size=None):
    """videoswin.py:86-99"""
    use_w = list(window_size)
    use_s = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_w[i] = x_size[i]
            if use_s is not None:
                use_s[i] = 0
    if shift_size is None:
        return tuple(use_w)
    return tuple(use_w), tuple(use_s)


def relative_position_index(window):
    """videoswin.py:113-127"""
    wd, wh, ww = window
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    flat = coords.flatten(1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


class _Linear(nn.Module):
    def __init__(self, in_f, out_f, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f).normal_(0, 0.02))
        self.bias = nn.Parameter(torch.zeros(out_f)) if bias else None


class _Norm(nn.Module):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = eps


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = _Linear(in_features, hidden_features)
        self.fc2 = _Linear(hidden_features, in_features)


class WindowAttention3D(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        n_rel = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_rel, num_heads).normal_(0, 0.02).clamp_(-0.04, 0.04))
        self.register_buffer("relative_position_index", relative_position_index(window_size))
        self.qkv = _Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = _Linear(dim, dim)


class SwinTransformerBlock3D(nn.Module):
    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_path=0.):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = _Norm(dim)
        self.attn = WindowAttention3D(dim, window_size, num_heads, qkv_bias, qk_scale)
        self.norm2 = _Norm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.drop_path_rate = drop_path

    def run(self, x, grid):
        """x: [B*D*H*W, C] natural order -> same.  (forward_part1/part2, videoswin.py:191-245)"""
        B, D, H, W = grid
        ws, ss = get_window_size((D, H, W), self.window_size, self.shift_size)
        if D % ws[0] or H % ws[1] or W % ws[2]:
            raise NotImplementedError(f"token grid {(D, H, W)} is not a multiple of window {ws}: the padded-window "
                                      "case (F=12, videoswin.py:196-203) is not built yet")
        hd = self.dim // self.num_heads
        geom = (grid, ws, ss, tuple(self.window_size), self.num_heads, hd, self.attn.scale)
        a = self.attn
        rng = getattr(self, "_rng", None) if self.training else None     # DropPath (videoswin.py:238,243): training only
        dp = self.drop_path_rate
        y, xr = Fn.layer_norm_residual(x, LN(self.norm1.weight, self.norm1.bias, self.norm1.eps))
        qkv = Fn.linear(y, lin_of(a.qkv.weight, a.qkv.bias))
        o = Fn.window_attention(qkv, a.relative_position_bias_table, geom)
        x = Fn.residual_branch(lambda r, s=None: Fn.linear(o, lin_of(a.proj.weight, a.proj.bias), residual=r, row_scale=s), xr, rng, dp, B)
        y, xr = Fn.layer_norm_residual(x, LN(self.norm2.weight, self.norm2.bias, self.norm2.eps))
        return Fn.residual_branch(
            lambda r, s=None: Fn.mlp(y, lin_of(self.mlp.fc1.weight, self.mlp.fc1.bias), lin_of(self.mlp.fc2.weight, self.mlp.fc2.bias),
                                     K.ACT_GELU, residual=r, row_scale=s), xr, rng, dp, B)


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.reduction = _Linear(4 * dim, 2 * dim, bias=False)
        self.norm = _Norm(4 * dim)

    def run(self, x, grid):
        B, D, H, W = grid
        y = Fn.PatchMergeFn.apply(x, B * D, H, W, self.dim)
        y = Fn.layer_norm(y, LN(self.norm.weight, self.norm.bias, self.norm.eps))
        return Fn.linear(y, lin_of(self.reduction.weight)), (B, D, H // 2, W // 2)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop_path=0., downsample=None):
        super().__init__()
        self.window_size = window_size
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth = depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim, num_heads, window_size, (0, 0, 0) if i % 2 == 0 else self.shift_size, mlp_ratio,
                                   qkv_bias, qk_scale, drop_path[i] if isinstance(drop_path, list) else drop_path)
            for i in range(depth)])
        self.downsample = downsample(dim=dim) if downsample is not None else None

    def run(self, x, grid, stage=0):
        bucket = None
        for i, blk in enumerate(self.blocks):
            if swin_bucket(stage, self.depth, i) != bucket:    # gradient all-reduce bucket boundary (distributed.py)
                bucket = swin_bucket(stage, self.depth, i)
                x = mark(x, bucket)
            x = blk.run(x, grid)
        if self.downsample is not None:
            x, grid = self.downsample.run(x, grid)
        return x, grid


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, time_stride=1):
        super().__init__()
        assert tuple(patch_size) == (2, 4, 4) and in_chans == 3 and time_stride == 1, "only the shipped patch geometry"
        self.patch_size, self.embed_dim = patch_size, embed_dim
        self.proj = nn.Module()
        self.proj.weight = nn.Parameter(torch.empty(embed_dim, in_chans, *patch_size).normal_(0, 0.02))
        self.proj.bias = nn.Parameter(torch.zeros(embed_dim))
        self.norm = _Norm(embed_dim)

    def run(self, video, anchor, dtype):
        """video [B,F,3,H,W] (the batch layout, modeling.py:451) -> tokens [B*F*(H/4)*(W/4), E]."""
        cols = K.swin_im2col(video, dtype)
        w = self.proj.weight
        lin = Fn.Lin(w.lp.view(self.embed_dim, -1), w.main_grad.view(self.embed_dim, -1), self.proj.bias.data,
                     self.proj.bias.main_grad)
        x = Fn.linear(cols, lin, anchor=anchor)
        return Fn.layer_norm(x, LN(self.norm.weight, self.norm.bias, self.norm.eps))


class SwinTransformer3D(nn.Module):
    """Constructor mirrors videoswin.py:379-399 (arguments the shipped configs never change keep
    their defaults and are validated)."""

    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(2, 4, 4), in_chans=3, embed_dim=96,
                 depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=(8, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None, patch_norm=True,
                 frozen_stages=-1, use_checkpoint=False, time_stride=1, checkpointing=False):
        super().__init__()
        assert patch_norm and drop_rate == 0. and attn_drop_rate == 0.
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.window_size = window_size
        self.patch_embed = PatchEmbed3D(patch_size, in_chans, embed_dim, time_stride)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(int(embed_dim * 2 ** i), depths[i], num_heads[i], window_size, mlp_ratio,
                                          qkv_bias, qk_scale, dpr[sum(depths[:i]):sum(depths[:i + 1])],
                                          PatchMerging if i < self.num_layers - 1 else None))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = _Norm(self.num_features)
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    def forward_tokens(self, video, dtype=None):
        """video [B,F,3,H,W] -> ([B*F*(H/32)*(W/32), 8E] tokens in (b,f,h,w) order, (B,F,H/32,W/32))."""
        B, F, _, H, W = video.shape
        dtype = dtype or self.norm.weight.lp.dtype
        anchor = self._anchor.requires_grad_(True) if torch.is_grad_enabled() else None
        x = self.patch_embed.run(video.contiguous(), anchor, dtype)
        grid = (B, F, H // 4, W // 4)
        for i, layer in enumerate(self.layers):
            x, grid = layer.run(x, grid, stage=i)
        x = Fn.layer_norm(x, LN(self.norm.weight, self.norm.bias, self.norm.eps))
        return x, grid

    def forward(self, x):
        """Reference signature (videoswin.py:441-458): [B,3,D,H,W] -> [B, 8E, D, H/32, W/32]."""
        tok, (B, D, H, W) = self.forward_tokens(x.transpose(1, 2).contiguous())
        return tok.view(B, D, H, W, -1).permute(0, 4, 1, 2, 3)
