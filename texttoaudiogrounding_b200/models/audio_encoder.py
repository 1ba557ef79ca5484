"""Cnn8Rnn audio encoder — host-side mirror of reference models/audio_encoder.py:89-232.

Same constructor kwargs, attributes (``embed_dim``, ``downsample_ratio``, ``time_resolution``),
dict-in / dict-out contract and state-dict keys as the reference; the forward and backward
passes run in the sm_100a kernels of libtag_b200.so (log-mel -> bn0 -> 4 x ConvBlock -> mean ->
fc1 -> BiGRU), scheduled by ``engine.encoder_forward`` / ``encoder_backward``.
The nn.BatchNorm2d / nn.Conv2d / nn.Linear / nn.GRU children are parameter holders only —
their own forward() is never called.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List

import torch
import torch.nn as nn

from .. import engine
from . import nn_ops
from ..frontend_consts import hann_window, slaney_mel_fbanks
from .base import LoadPretrainedMixin, merge_matched_keys


def default_compute_dtype() -> torch.dtype:
    return {"fp32": torch.float32, "bf16": torch.bfloat16}[os.environ.get("TAG_B200_PRECISION", "bf16")]


def init_layer(layer):
    """Xavier-uniform weight, zero bias (reference models/panns.py:5-11)."""
    nn.init.xavier_uniform_(layer.weight)
    if hasattr(layer, "bias") and layer.bias is not None:
        layer.bias.data.fill_(0.)


def init_bn(bn):
    """reference models/panns.py:14-17"""
    bn.bias.data.fill_(0.)
    bn.weight.data.fill_(1.)


class ConvBlock(nn.Module):
    """Parameter holder with the reference's names (models/panns.py:20-45)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=(1, 1),
                               padding=(1, 1), bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1),
                               padding=(1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        init_layer(self.conv1)
        init_layer(self.conv2)
        init_bn(self.bn1)
        init_bn(self.bn2)
        for conv in (self.conv1, self.conv2):      # kernels read [Cout][kh][kw][Cin]
            conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)

    def forward(self, *a, **k):
        raise RuntimeError("ConvBlock is a parameter holder; the fused CUDA path in Cnn8Rnn.forward "
                           "is the only implementation (no PyTorch fallback)")


class _Buffers(nn.Module):
    def __init__(self, name: str, value: torch.Tensor):
        super().__init__()
        self.register_buffer(name, value)


class _MelSpecBuffers(nn.Module):
    """Holds ``spectrogram.window`` and ``mel_scale.fb`` under torchaudio's buffer names."""

    def __init__(self, sample_rate: int, n_fft: int, f_max: float):
        super().__init__()
        self.spectrogram = _Buffers("window", hann_window(n_fft))
        self.mel_scale = _Buffers("fb", slaney_mel_fbanks(n_fft // 2 + 1, 50.0, float(f_max), 64,
                                                           sample_rate))


def _packed_conv(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,3,3] parameter -> tensor whose memory is [Cout][kh][kw][Cin] (zero-copy when the
    parameter is channels_last)."""
    v = w.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _cat_if_needed(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.is_contiguous() and b.is_contiguous() and \
            b.data_ptr() == a.data_ptr() + a.numel() * a.element_size():
        return torch.as_strided(a, (a.shape[0] + b.shape[0],) + tuple(a.shape[1:]), a.stride())
    return torch.cat([a, b], dim=0)


class _Cnn8RnnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, waveform, *params):
        training = module.training
        Wt = module._weights()
        # needs_input_grad ignores torch.no_grad(); the module records the grad mode before apply()
        save = module._grad_mode and any(ctx.needs_input_grad)
        module._call_count += 1
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * module._call_count) & 0x7FFFFFFFFFFFFFFF
        out, ectx = engine.encoder_forward(
            Wt, waveform, training=training, bn_training=module.bn0.training,
            dropout=module.dropout_enabled, seed=seed, dtype=module.compute_dtype, save=save,
            seed_dev=nn_ops.SEED_DEV, stages=module._stages)
        ctx.module, ctx.ectx, ctx.Wt = module, ectx, Wt
        return out

    @staticmethod
    def backward(ctx, d_out):
        module, ectx, Wt = ctx.module, ctx.ectx, ctx.Wt
        if ectx is None:
            raise RuntimeError("backward through Cnn8Rnn without saved activations")
        G = module._zero_grads(Wt)
        engine.encoder_backward(Wt, ectx, d_out, G)
        return (None, None) + tuple(module._grads_to_param_order(G))


class Cnn8Rnn(nn.Module, LoadPretrainedMixin):
    def __init__(self, sample_rate: int, freeze_cnn: bool = False, freeze_bn: bool = False,
                 pretrained: "str | None" = None, output_fn: callable = sys.stdout.write,
                 compute_dtype: "str | torch.dtype | None" = None):
        super().__init__()
        self.downsample_ratio = 4
        self.time_resolution = 0.04
        self.freeze_cnn = freeze_cnn
        self.freeze_bn = freeze_bn
        self.hop_length = int(0.010 * sample_rate)
        self.win_length = int(0.032 * sample_rate)
        if self.win_length != 1024 or self.hop_length != 320:
            raise ValueError("the CUDA frontend is built for sample_rate=32000 (n_fft 1024, hop 320)")
        f_max = 14000 if sample_rate == 32000 else int(sample_rate / 2)
        self.melspec_extractor = _MelSpecBuffers(sample_rate, self.win_length, f_max)

        self.bn0 = nn.BatchNorm2d(64)
        self.conv_block1 = ConvBlock(in_channels=1, out_channels=64)
        self.conv_block2 = ConvBlock(in_channels=64, out_channels=128)
        self.conv_block3 = ConvBlock(in_channels=128, out_channels=256)
        self.conv_block4 = ConvBlock(in_channels=256, out_channels=512)
        self.fc1 = nn.Linear(512, 512, bias=True)
        self.rnn = nn.GRU(512, 256, bidirectional=True, batch_first=True)
        self.embed_dim = 512
        self.init_weight()

        if isinstance(compute_dtype, str):
            compute_dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[compute_dtype]
        self.compute_dtype = compute_dtype or default_compute_dtype()
        self.dropout_enabled = True       # parity tests switch the five F.dropout sites off
        self._stages = None               # set to a dict to capture stage tensors (tests)
        self._call_count = 0
        self._grad_mode = True
        self._mel_range = None
        self._mel_nnz = 0

        if pretrained is not None:
            self.load_pretrained(pretrained, output_fn)
        if self.freeze_cnn:
            for param in self.parameters():
                param.requires_grad = False
            for param in self.rnn.parameters():
                param.requires_grad = True

    # ------------------------------------------------------------------ reference surface
    def process_state_dict(self, model_dict, pretrained_dict, output_fn, model_name):
        if "model" in pretrained_dict:          # the mixin may already have unwrapped the checkpoint
            pretrained_dict = pretrained_dict["model"]
        return merge_matched_keys(model_dict, pretrained_dict, output_fn, model_name)

    def train(self, mode: bool = True):
        super().train(mode=mode)
        if self.freeze_bn:
            for module in self.modules():
                if module.__class__.__name__.find("BatchNorm") != -1:
                    module.eval()
        return self

    def init_weight(self):
        init_bn(self.bn0)
        init_layer(self.fc1)

    def _load_from_state_dict(self, *args, **kwargs):
        self._mel_range = None
        self._mel_nnz = 0
        return super()._load_from_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ kernel plumbing
    def _bns(self) -> List[nn.BatchNorm2d]:
        bns = [self.bn0]
        for blk in (self.conv_block1, self.conv_block2, self.conv_block3, self.conv_block4):
            bns += [blk.bn1, blk.bn2]
        return bns

    def _convs(self) -> List[nn.Conv2d]:
        convs = []
        for blk in (self.conv_block1, self.conv_block2, self.conv_block3, self.conv_block4):
            convs += [blk.conv1, blk.conv2]
        return convs

    def _param_list(self) -> List[nn.Parameter]:
        """Fixed order used by the autograd function."""
        ps = []
        for bn in self._bns():
            ps += [bn.weight, bn.bias]
        ps += [c.weight for c in self._convs()]
        ps += [self.fc1.weight, self.fc1.bias]
        r = self.rnn
        ps += [r.weight_ih_l0, r.weight_ih_l0_reverse, r.bias_ih_l0, r.bias_ih_l0_reverse,
               r.weight_hh_l0, r.weight_hh_l0_reverse, r.bias_hh_l0, r.bias_hh_l0_reverse]
        return ps

    def _weights(self) -> engine.EncoderWeights:
        fb = self.melspec_extractor.mel_scale.fb
        if self._mel_range is None or self._mel_range.device != fb.device:
            self._mel_range = engine.compute_mel_range(fb)
            self._mel_nnz = engine.mel_nnz(self._mel_range)
        r = self.rnn
        return engine.EncoderWeights(
            window=self.melspec_extractor.spectrogram.window, fb=fb, mel_range=self._mel_range, mel_nnz=self._mel_nnz,
            bn=[(bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var) for bn in self._bns()],
            conv=[_packed_conv(c.weight.data) for c in self._convs()],
            fc_w=self.fc1.weight.data, fc_b=self.fc1.bias.data,
            w_ih=_cat_if_needed(r.weight_ih_l0.data, r.weight_ih_l0_reverse.data),
            b_ih=_cat_if_needed(r.bias_ih_l0.data, r.bias_ih_l0_reverse.data),
            w_hh=_cat_if_needed(r.weight_hh_l0.data, r.weight_hh_l0_reverse.data).view(2, 768, 256),
            b_hh=_cat_if_needed(r.bias_hh_l0.data, r.bias_hh_l0_reverse.data))

    def _zero_grads(self, Wt: engine.EncoderWeights) -> engine.EncoderGrads:
        z = torch.zeros_like
        return engine.EncoderGrads(
            bn=[(z(g), z(b)) for g, b, _, _ in Wt.bn], conv=[z(c) for c in Wt.conv],
            fc_w=z(Wt.fc_w), fc_b=z(Wt.fc_b), w_ih=z(Wt.w_ih), b_ih=z(Wt.b_ih),
            w_hh=z(Wt.w_hh), b_hh=z(Wt.b_hh))

    @staticmethod
    def _grads_to_param_order(G: engine.EncoderGrads):
        out = []
        for dg, db in G.bn:
            out += [dg, db]
        out += [g.permute(0, 3, 1, 2) for g in G.conv]        # [Co][kh][kw][Ci] -> [Co,Ci,3,3] view
        out += [G.fc_w, G.fc_b]
        out += [G.w_ih[:768], G.w_ih[768:], G.b_ih[:768], G.b_ih[768:],
                G.w_hh[0], G.w_hh[1], G.b_hh[:768], G.b_hh[768:]]
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, input_dict: Dict):
        """Input: waveform (batch_size, n_samples) -> {"embedding": [B, T', 512], "length": [B]}"""
        waveform = input_dict["waveform"]
        specaug = input_dict["specaug"]
        if self.training and specaug:
            raise NotImplementedError("SpecAugment is outside the B200 hot path (SURVEY.md §8a a4: "
                                      "run_strong.py always passes specaug=False)")
        mixup_lambda = input_dict.get("mixup_lambda", None)
        if self.training and mixup_lambda is not None:
            raise NotImplementedError("mixup is outside the B200 hot path")
        if not waveform.is_cuda:
            raise RuntimeError("Cnn8Rnn (B200) needs CUDA tensors: there is no CPU fallback")
        waveform = waveform.float()
        self._grad_mode = torch.is_grad_enabled()
        x = _Cnn8RnnFunction.apply(self, waveform, *self._param_list())
        if self.training:
            for bn in self._bns():
                if bn.training and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1

        length = torch.div(torch.as_tensor(input_dict["waveform_len"]), self.hop_length,
                           rounding_mode="floor") + 1
        length = torch.div(length, self.downsample_ratio, rounding_mode="floor")
        return {"embedding": x, "length": length}


# the reference's example configs name this class ``models.audio_encoder.Cnn8_Rnn`` (eg_configs/weakly_supervised/**),
# a spelling its own module no longer defines; the alias lets those YAMLs resolve
Cnn8_Rnn = Cnn8Rnn
