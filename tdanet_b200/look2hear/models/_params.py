"""Parameter containers: the module tree that owns TDANet's weights under the reference's
state_dict keys and default initialisation.  These modules never compute; `forward` of the model
classes hands their tensors to the CUDA engine.

Key names / shapes / construction order follow the reference so that checkpoints interchange and
`torch.manual_seed(s)` gives identical random-init weights:
  TDANet_best.py:302-340 (UConvBlock), :383-390 (Recurrent), :429-462 (front/back end);
  TDANet.py:509-579 (fork UConvBlock with conv_pool); TDANet_mult_tes.py:317-342 (ConvEncoder).
"""
import math

import torch
import torch.nn as nn


class GlobLN(nn.Module):
    """gamma/beta holder of the hand-written GlobLN (TDANet_best.py:33-64)."""

    def __init__(self, channels):
        super().__init__()
        self.channel_size = channels
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


def make_gln(variant, channels):
    # the fork / multi-resolution files define GlobLN as GroupNorm(1, C, eps=1e-8): keys weight/bias
    return GlobLN(channels) if variant == "best" else nn.GroupNorm(1, channels, eps=1e-8)


class ConvNorm(nn.Module):
    def __init__(self, variant, n_in, n_out, k, stride=1, groups=1, bias=True, act=False):
        super().__init__()
        self.conv = nn.Conv1d(n_in, n_out, k, stride=stride, padding=(k - 1) // 2, bias=bias, groups=groups)
        self.norm = make_gln(variant, n_out)
        if act:
            self.act = nn.PReLU()


class SepConvNorm(nn.Module):
    def __init__(self, variant, ch, k, stride):
        super().__init__()
        self.dw_conv = nn.Conv1d(ch, ch, k, stride=stride, padding=(k - 1) // 2, groups=ch)
        self.pw_conv = nn.Conv1d(ch, ch, 1)
        self.norm = make_gln(variant, ch)


class LA(nn.Module):
    def __init__(self, variant, ch, k):
        super().__init__()
        self.local_embedding = ConvNorm(variant, ch, ch, k, groups=ch, bias=False)
        self.global_embedding = ConvNorm(variant, ch, ch, k, groups=ch, bias=False)
        self.global_act = ConvNorm(variant, ch, ch, k, groups=ch, bias=False)


class PositionalEncoding(nn.Module):
    def __init__(self, channels, max_length):
        super().__init__()
        pos = torch.arange(0, max_length).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, channels, 2, dtype=torch.float) * -(math.log(10000.0) / channels))
        pe = torch.zeros(max_length, channels)
        pe[:, 0::2] = torch.sin(pos * freq)
        pe[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", pe.unsqueeze(0))


class AttentionParams(nn.Module):
    def __init__(self, channels, n_head, dropout, batch_first):
        super().__init__()
        self.pos_enc = PositionalEncoding(channels, 10000)
        self.attn_in_norm = nn.LayerNorm(channels)
        self.attn = nn.MultiheadAttention(channels, n_head, dropout, batch_first=batch_first)
        self.norm = nn.LayerNorm(channels)


class FFNParams(nn.Module):
    def __init__(self, variant, channels, hidden):
        super().__init__()
        self.fc1 = ConvNorm(variant, channels, hidden, 1, bias=False)
        self.dwconv = nn.Conv1d(hidden, hidden, 5, 1, 2, bias=True, groups=hidden)
        self.fc2 = ConvNorm(variant, hidden, channels, 1, bias=False)


class GAParams(nn.Module):
    def __init__(self, variant, channels, n_head):
        super().__init__()
        self.attn = AttentionParams(channels, n_head, 0.1, batch_first=(variant == "multres"))
        self.mlp = FFNParams(variant, channels, channels * 2)


class UConvBlockParams(nn.Module):
    def __init__(self, variant, out_channels, in_channels, depth, n_head):
        super().__init__()
        C = in_channels
        self.proj_1x1 = ConvNorm(variant, out_channels, C, 1, act=True)
        self.spp_dw = nn.ModuleList([ConvNorm(variant, C, C, 5, stride=1, groups=C)])
        if variant == "fork":
            self.conv_pool = nn.ModuleList([SepConvNorm(variant, C, 5, 1)])
        for i in range(1, depth):
            self.spp_dw.append(ConvNorm(variant, C, C, 5, stride=2, groups=C))
            if variant == "fork":
                s = 2 ** i
                self.conv_pool.append(SepConvNorm(variant, C, 2 * s + 1, s))
        if variant == "best":
            self.loc_glo_fus = nn.ModuleList([LA(variant, C, 1) for _ in range(depth)])
        self.res_conv = nn.Conv1d(C, out_channels, 1)
        self.globalatt = GAParams(variant, C, n_head)
        self.last_layer = nn.ModuleList([LA(variant, C, 5) for _ in range(depth - 1)])


class RecurrentParams(nn.Module):
    def __init__(self, variant, out_channels, in_channels, depth, n_head):
        super().__init__()
        self.unet = UConvBlockParams(variant, out_channels, in_channels, depth, n_head)
        self.concat_block = nn.Sequential(
            nn.Conv1d(out_channels, out_channels, 1, 1, groups=out_channels), nn.PReLU())


class ConvEncoderParams(nn.Module):
    def __init__(self, base_ks, out_channels, kernels):
        super().__init__()
        self.conv_list = nn.ModuleList(
            nn.Conv1d(1, out_channels // kernels, k * base_ks, stride=base_ks // 4, padding=k * base_ks // 2, bias=False)
            for k in range(1, kernels + 1))
