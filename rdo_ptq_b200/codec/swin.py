"""Swin-transformer pieces of the Lu2022 codec (task-oriented-PTQ/models/layers.py: WindowAttention :86-183,
SwinTransformerBlock :186-306, BasicLayer :309-366, RSTB :369-433) over libb200lic, FORWARD ONLY.

Parameter and buffer names follow the reference so that its state dicts load.  Every Linear runs on the tcgen05 conv engine
(ops.linear), LayerNorm / GELU / the attention core (ops.window_attn_softmax, ops.window_attn_apply) are libb200lic kernels;
cyclic shift and window (un)partition are index permutations of whole tensors and stay torch views + copies (buffer
plumbing, like pad / crop in evaluate.py).  Dropout and stochastic depth are the identity at evaluation and are not carried.
"""
import torch
import torch.nn as nn

from .. import ops
from .layers import Mlp


def to_windows(x, ws):
    """[B, H, W, C] -> [B * H/ws * W/ws, ws*ws, C] (row-major windows)."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def from_windows(w, ws, H, W):
    """Inverse of `to_windows`: [B * H/ws * W/ws, ws*ws, C] -> [B, H, W, C]."""
    C = w.shape[-1]
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.reshape(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H, W, C)


def shift_mask(H, W, ws, shift):
    """[nW, ws*ws, ws*ws] additive mask of shifted-window attention: -100 between tokens that come from different sides of
    the cyclic shift, 0 otherwise (layers.py:229-251)."""
    region = torch.zeros(H, W)
    cuts = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    k = 0
    for hs in cuts:
        for wsl in cuts:
            region[hs, wsl] = k
            k += 1
    ids = to_windows(region.reshape(1, H, W, 1), ws).reshape(-1, ws * ws)
    diff = ids.unsqueeze(1) - ids.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def relative_position_index(ws_h, ws_w):
    """[N, N] index into the (2*ws_h-1)*(2*ws_w-1) bias table for every (query, key) pair of a window."""
    ii, jj = torch.meshgrid(torch.arange(ws_h), torch.arange(ws_w), indexing="ij")
    pos = torch.stack([ii.reshape(-1), jj.reshape(-1)])                    # [2, N]
    rel = pos[:, :, None] - pos[:, None, :]                                 # [2, N, N]
    return (rel[0] + ws_h - 1) * (2 * ws_w - 1) + (rel[1] + ws_w - 1)


def gathered_bias(table, index, n_tokens):
    """relative_position_bias_table [T, nH] + index [N, N] -> [nH, N, N] contiguous (layers.py:150-153)."""
    return table[index.reshape(-1)].reshape(n_tokens, n_tokens, -1).permute(2, 0, 1).contiguous()


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(wh, ww))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)

    def forward(self, x, mask=None):
        n = x.shape[1]
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias)
        bias = gathered_bias(self.relative_position_bias_table, self.relative_position_index, n)
        attn = ops.window_attn_softmax(qkv, bias, mask, self.num_heads, self.scale)
        return ops.linear(ops.window_attn_apply(attn, qkv), self.proj.weight, self.proj.bias)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=None, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:      # one window covers the map: no partition, no shift
            self.shift_size, self.window_size = 0, min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, (self.window_size, self.window_size), num_heads, qkv_bias, qk_scale)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio)) if act_layer is None else Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer)
        self.register_buffer("attn_mask", shift_mask(*self.input_resolution, self.window_size, self.shift_size)
                             if self.shift_size > 0 else None)

    def mask_for(self, x_size, device):
        if self.shift_size == 0:
            return None
        if tuple(x_size) == self.input_resolution:
            return self.attn_mask
        return shift_mask(x_size[0], x_size[1], self.window_size, self.shift_size).to(device)

    def attend(self, attn, x, x_size):
        """norm'ed tokens [B, L, C] -> attention output [B, L, C]: shift, partition, `attn`, merge, shift back."""
        H, W = x_size
        B, _, C = x.shape
        x = x.reshape(B, H, W, C)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        win = attn(to_windows(x, self.window_size).contiguous(), mask=self.mask_for(x_size, x.device))
        x = from_windows(win, self.window_size, H, W)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        return x.reshape(B, H * W, C)

    def forward(self, x, x_size):
        y = self.attend(self.attn, ops.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps), x_size)
        x = ops.add_act(x, y.contiguous())
        return ops.add_act(x, self.mlp(ops.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)))


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, input_resolution, depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2,
                                 mlp_ratio, qkv_bias, qk_scale, norm_layer=norm_layer) for i in range(depth)])

    def forward(self, x, x_size):
        for blk in self.blocks:
            x = blk(x, x_size)
        return x


class PatchEmbed(nn.Module):
    def forward(self, x):                       # [B, C, H, W] -> [B, H*W, C]
        return x.flatten(2).transpose(1, 2).contiguous()


class PatchUnEmbed(nn.Module):
    def forward(self, x, x_size):               # [B, H*W, C] -> [B, C, H, W]
        B, _, C = x.shape
        return x.transpose(1, 2).reshape(B, C, x_size[0], x_size[1]).contiguous()


class RSTB(nn.Module):
    """Residual Swin transformer block: tokens through a BasicLayer, back to a feature map, plus the input."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution = dim, input_resolution
        self.residual_group = BasicLayer(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale,
                                         norm_layer=norm_layer, use_checkpoint=use_checkpoint)
        self.patch_embed, self.patch_unembed = PatchEmbed(), PatchUnEmbed()

    def forward(self, x, x_size):
        y = self.patch_unembed(self.residual_group(self.patch_embed(x), x_size), x_size)
        return ops.add_act(y, x)
