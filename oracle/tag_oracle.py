"""CPU oracle for the cnn8rnn-w2vmean hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 PyTorch-on-CPU ops, the arithmetic the reference
(wsntxxn/TextToAudioGrounding) performs on its strong-supervision hot path.  It is the
checker the CUDA path is compared against; it is *never* imported by the product
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity pin: the reference holds no tests or golden vectors, so this
oracle is pinned against outputs of the reference itself, imported read-only in the
build container by ``oracle/make_golden.py`` and committed as ``tests/golden/*.npz``
(see ``tests/test_oracle_golden.py``).

Reference sites restated here (all relative to /root/reference):
  * log-mel frontend ............ models/audio_encoder.py:113-124,183-184 (torchaudio
                                  MelSpectrogram + AmplitudeToDB; algorithm lives in
                                  torchaudio 2.11.0 `functional.spectrogram`,
                                  `functional.melscale_fbanks`, `amplitude_to_DB`)
  * bn0 over mel bins ........... models/audio_encoder.py:188-190
  * ConvBlock x4 + dropout ...... models/panns.py:47-62, models/audio_encoder.py:202-211
  * freq mean, fc1, BiGRU ....... models/audio_encoder.py:212-217
  * length arithmetic ........... models/audio_encoder.py:219-227
  * EmbeddingAgg (mean) ......... models/text_encoder.py:39-43,79-88; models/utils.py:33-58
  * BiEncoder orchestration ..... models/audio_text_model.py:58-98
  * DotProduct match ............ models/match.py:43-60
  * Runner.forward post-step .... python_scripts/training/run_strong.py:92-120
  * FrameBceLoss ................ losses.py:12-24; models/utils.py:22-30
  * train step .................. python_scripts/training/run_strong.py:139-147
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

N_FFT = 1024
HOP = 320
N_MELS = 64
SAMPLE_RATE = 32000
F_MIN = 50.0
F_MAX = 14000.0
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
VOCAB = 5221
EMBED = 512
HIDDEN = 256

CONV_CHANNELS = [(1, 64), (64, 128), (128, 256), (256, 512)]
POOLS = [(2, 2), (2, 2), (1, 2), (1, 2)]


# --------------------------------------------------------------------------- frontend
def _hz_to_mel_slaney(f: float) -> float:
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        mel = min_log_mel + math.log(f / min_log_hz) / logstep
    return mel


def _mel_to_hz_slaney(mels: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def melscale_fbanks(n_freqs: int = N_FFT // 2 + 1, f_min: float = F_MIN,
                    f_max: float = F_MAX, n_mels: int = N_MELS,
                    sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """Slaney-scale, slaney-normalised triangular filterbank [n_freqs, n_mels]
    (torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney"))."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_slaney(f_min)
    m_max = _hz_to_mel_slaney(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0)
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def hann_window(n: int = N_FFT) -> torch.Tensor:
    """Periodic Hann (torch.hann_window default)."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * k / n)).float()


def power_spectrogram(waveform: torch.Tensor, window: torch.Tensor) -> torch.Tensor:
    """[B, L] -> [B, 513, T0]; center=True, reflect pad, |rfft|^2 (power=2)."""
    pad = N_FFT // 2
    x = F.pad(waveform.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(-1, N_FFT, HOP)                      # [B, T0, 1024]
    spec = torch.fft.rfft(frames * window, dim=-1)         # [B, T0, 513]
    return (spec.real ** 2 + spec.imag ** 2).transpose(1, 2)


def logmel_db(waveform: torch.Tensor, window: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """[B, L] -> dB log-mel [B, 64, T0]  (MelScale matmul + AmplitudeToDB(power))."""
    power = power_spectrogram(waveform, window)            # [B, 513, T0]
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
    return 10.0 * torch.log10(torch.clamp(mel, min=1e-10))


# --------------------------------------------------------------------------- weights
def state_dict_spec(vocab: int = VOCAB):
    """(key, shape, kind) for every entry of the reference BiEncoder state dict on the
    cnn8rnn-w2vmean config (SURVEY.md §8b 'State-dict keys')."""
    spec = []
    ae = "audio_encoder."
    spec.append((ae + "melspec_extractor.spectrogram.window", (N_FFT,), "window"))
    spec.append((ae + "melspec_extractor.mel_scale.fb", (N_FFT // 2 + 1, N_MELS), "fb"))
    def bn(prefix, c):
        spec.append((prefix + ".weight", (c,), "bn_w"))
        spec.append((prefix + ".bias", (c,), "bn_b"))
        spec.append((prefix + ".running_mean", (c,), "bn_rm"))
        spec.append((prefix + ".running_var", (c,), "bn_rv"))
        spec.append((prefix + ".num_batches_tracked", (), "bn_n"))
    bn(ae + "bn0", N_MELS)
    for i, (ci, co) in enumerate(CONV_CHANNELS, 1):
        p = f"{ae}conv_block{i}."
        spec.append((p + "conv1.weight", (co, ci, 3, 3), "xavier"))
        spec.append((p + "conv2.weight", (co, co, 3, 3), "xavier"))
        bn(p + "bn1", co)
        bn(p + "bn2", co)
    spec.append((ae + "fc1.weight", (EMBED, EMBED), "xavier"))
    spec.append((ae + "fc1.bias", (EMBED,), "bias"))
    for sfx in ("", "_reverse"):
        spec.append((f"{ae}rnn.weight_ih_l0{sfx}", (3 * HIDDEN, EMBED), "gru"))
        spec.append((f"{ae}rnn.weight_hh_l0{sfx}", (3 * HIDDEN, HIDDEN), "gru"))
        spec.append((f"{ae}rnn.bias_ih_l0{sfx}", (3 * HIDDEN,), "gru"))
        spec.append((f"{ae}rnn.bias_hh_l0{sfx}", (3 * HIDDEN,), "gru"))
    spec.append(("text_encoder.embedding.core.weight", (vocab, EMBED), "emb"))
    return spec


def synth_state_dict(seed: int = 1, vocab: int = VOCAB, sharpen: float = 1.0,
                     perturb_bn: bool = False) -> Dict[str, torch.Tensor]:
    """Implementation-independent deterministic weights: every tensor is drawn from its
    own generator keyed by (seed, position in spec), with the reference's init
    distributions (models/panns.py:5-17 Xavier-uniform; nn.GRU U(+-1/sqrt(H));
    models/utils.py:18-19 Kaiming-uniform embedding).  ``sharpen`` scales the embedding
    so logits leave the +-0.01 band of the raw init; ``perturb_bn``
    gives BN affine/running stats non-trivial values so eval-mode parity is not vacuous."""
    sd = {}
    for idx, (key, shape, kind) in enumerate(state_dict_spec(vocab)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if kind == "window":
            t = hann_window()
        elif kind == "fb":
            t = melscale_fbanks()
        elif kind == "xavier":
            if len(shape) == 4:
                fan_out, fan_in = shape[0] * 9, shape[1] * 9
            else:
                fan_out, fan_in = shape
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "bias":
            t = torch.zeros(shape)
            if perturb_bn:
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "gru":
            a = 1.0 / math.sqrt(HIDDEN)
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "emb":
            a = math.sqrt(6.0 / EMBED)
            t = (torch.rand(shape, generator=g) * 2 - 1) * a * sharpen
        elif kind == "bn_w":
            t = torch.ones(shape)
            if perturb_bn:
                t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            t = torch.zeros(shape)
            if perturb_bn:
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        elif kind == "bn_rm":
            t = torch.zeros(shape)
            if perturb_bn:
                if key.endswith("bn0.running_mean"):
                    t = -10.0 + torch.rand(shape, generator=g) * 2
                else:
                    t = (torch.rand(shape, generator=g) * 2 - 1) * 0.3
        elif kind == "bn_rv":
            t = torch.ones(shape)
            if perturb_bn:
                if key.endswith("bn0.running_var"):
                    t = 5.0 + 4.0 * torch.rand(shape, generator=g)
                else:
                    t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "bn_n":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        sd[key] = t
    return sd


def synth_batch(batch: int, n_samples: int, n_tokens: int = 8, seed: int = 0,
                vocab: int = VOCAB, ragged: bool = False, tonal: bool = True):
    """Seeded synthetic batch in the collate schema of datasets/collate_function.py:43-84
    (SURVEY.md §8d 'Synthetic inputs')."""
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(batch, n_samples, generator=g)
    if tonal:
        t = torch.arange(n_samples, dtype=torch.float32) / SAMPLE_RATE
        for b in range(batch):
            for k in range(3):
                f = 110.0 * (b + 1) * (2 ** k) + 37.0 * k
                wav[b] += 0.05 * torch.sin(2 * math.pi * f * t)
    wav_len = [n_samples] * batch
    text_len = [n_tokens] * batch
    if ragged:
        fr = [1.0, 0.9, 0.75, 0.5]
        wav_len = [int(n_samples * fr[b % 4]) for b in range(batch)]
        for b in range(batch):
            wav[b, wav_len[b]:] = 0.0
        text_len = [3 + (b * 5) % (n_tokens - 2) for b in range(batch)]
        text_len[0] = n_tokens
    text = torch.randint(2, vocab, (batch, n_tokens), generator=g)
    for b in range(batch):
        text[b, text_len[b]:] = 0
    t_out = (n_samples // HOP + 1) // 4
    label = (torch.rand(batch, t_out + 1, generator=g) > 0.5).float()
    import numpy as np
    return {
        "waveform": wav,
        "waveform_len": np.asarray(wav_len, dtype=np.int64),
        "text": text,
        "text_len": torch.as_tensor(text_len, dtype=torch.long),
        "label": label,
    }


# --------------------------------------------------------------------------- model
def _bn(x, sd, prefix, training, stages=None):
    """BatchNorm2d forward; in training mode also updates running stats in ``sd``
    (momentum 0.1, unbiased var) exactly as nn.BatchNorm2d does."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        with torch.no_grad():
            sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm if not training else rm, rv, sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One direction of nn.GRU, h0 = 0, gate order [r; z; n]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi_all = F.linear(x, w_ih, b_ih)                       # [B, T, 3H]
    h = x.new_zeros(B, H)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[:, t]
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def bigru(x, sd, prefix="audio_encoder.rnn.", fast: bool = False):
    if fast:  # torch's fused CPU GRU (what nn.GRU calls); used for baseline timing
        flat = [sd[prefix + n] for n in (
            "weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
            "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse",
            "bias_hh_l0_reverse")]
        h0 = x.new_zeros(2, x.shape[0], HIDDEN)
        out, _ = torch._VF.gru(x, h0, flat, True, 1, 0.0, False, True, True)
        return out
    fw = gru_direction(x, sd[prefix + "weight_ih_l0"], sd[prefix + "weight_hh_l0"],
                       sd[prefix + "bias_ih_l0"], sd[prefix + "bias_hh_l0"], False)
    bw = gru_direction(x, sd[prefix + "weight_ih_l0_reverse"],
                       sd[prefix + "weight_hh_l0_reverse"],
                       sd[prefix + "bias_ih_l0_reverse"],
                       sd[prefix + "bias_hh_l0_reverse"], True)
    return torch.cat([fw, bw], dim=-1)


def _dropout(x, p, training, masks, name):
    if not training or p == 0.0:
        return x
    if masks is not None:
        m = masks.get(name)
        if m is None:
            return x
        return x * m            # mask already carries the 1/(1-p) scale
    return F.dropout(x, p=p, training=True)


def cnn8rnn_forward(sd, waveform, waveform_len, training: bool = False,
                    dropout_masks: Optional[dict] = None, dropout: bool = True,
                    stages: Optional[dict] = None, fast_gru: bool = False):
    """models/audio_encoder.py:178-232 with specaug=False and no mixup.

    ``dropout_masks`` (name -> multiplicative mask in NCHW / [B,T,C] layout) injects
    explicit masks; ``dropout=False`` disables dropout in training mode (parity runs)."""
    ae = "audio_encoder."
    x = logmel_db(waveform, sd[ae + "melspec_extractor.spectrogram.window"],
                  sd[ae + "melspec_extractor.mel_scale.fb"])          # [B, 64, T0]
    if stages is not None:
        stages["logmel_db"] = x
    x = x.transpose(1, 2).unsqueeze(1)                                  # [B,1,T0,64]
    x = x.transpose(1, 3)
    x = _bn(x, sd, ae + "bn0", training)
    if stages is not None:
        stages["bn0"] = x            # [B, 64, T0, 1], the layout nn.BatchNorm2d sees
    x = x.transpose(1, 3)
    use_do = training and dropout
    for i in range(4):
        p = f"{ae}conv_block{i + 1}."
        x = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], padding=1), sd, p + "bn1", training))
        x = F.relu(_bn(F.conv2d(x, sd[p + "conv2.weight"], padding=1), sd, p + "bn2", training))
        x = F.avg_pool2d(x, kernel_size=POOLS[i]) + F.max_pool2d(x, kernel_size=POOLS[i])
        x = _dropout(x, 0.2, use_do, dropout_masks, f"block{i + 1}")
        if stages is not None:
            stages[f"conv_block{i + 1}"] = x
    x = torch.mean(x, dim=3).transpose(1, 2)                            # [B, T', 512]
    x = _dropout(x, 0.5, use_do, dropout_masks, "fc_in")
    x = F.relu(F.linear(x, sd[ae + "fc1.weight"], sd[ae + "fc1.bias"]))
    if stages is not None:
        stages["fc1"] = x
    x = bigru(x, sd, ae + "rnn.", fast=fast_gru)
    if stages is not None:
        stages["rnn"] = x
    length = torch.div(torch.as_tensor(waveform_len), HOP, rounding_mode="floor") + 1
    length = torch.div(length, 4, rounding_mode="floor")
    return {"embedding": x, "length": length}


def generate_length_mask(lens, max_length=None):
    lens = torch.as_tensor(lens)
    if max_length is None:
        max_length = int(lens.max().item())
    idx = torch.arange(max_length).unsqueeze(0)
    return idx < lens.view(-1, 1)


def embedding_mean(sd, text, text_len):
    """models/text_encoder.py:39-43,79-88 + models/utils.py:33-58 (aggregation='mean')."""
    emb = F.embedding(text.long(), sd["text_encoder.embedding.core.weight"])   # [B,N,D]
    lens = torch.as_tensor(text_len)
    mask = generate_length_mask(lens, emb.size(1)).unsqueeze(-1)
    seq = (emb * mask).sum(1) / lens.view(-1, 1).to(emb.dtype)
    return {"token_emb": emb, "seq_emb": seq}


def dot_product_match(audio_emb, seq_emb, scale: bool = True):
    """models/match.py:43-60 (l2norm=False, text_level='seq'). Returns (frame_sim, logits)."""
    score = (audio_emb * seq_emb.unsqueeze(1)).sum(-1)
    if scale:
        score = score / math.sqrt(audio_emb.size(-1))
    return torch.sigmoid(score).clamp(1e-7, 1.0), score


def biencoder_forward(sd, input_dict, training=False, dropout_masks=None, dropout=True,
                      stages=None, fast_gru=False):
    """models/audio_text_model.py:58-98 (no cross encoder / projections / upsample)."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, stages, fast_gru)
    t = embedding_mean(sd, input_dict["text"], input_dict["text_len"])
    frame_sim, logits = dot_product_match(a["embedding"], t["seq_emb"])
    if stages is not None:
        stages["seq_emb"] = t["seq_emb"]
        stages["logits"] = logits
    return {"frame_sim": frame_sim, "length": a["length"]}


def runner_forward(sd, batch, training=True, **kw):
    """python_scripts/training/run_strong.py:92-120."""
    out = biencoder_forward(sd, batch, training=training, **kw)
    if training:
        label = batch["label"].float()
        frame_sim = out["frame_sim"]
        trunc = min(frame_sim.size(1), label.size(1))
        out.update({
            "frame_sim": frame_sim[..., :trunc],
            "label": label[..., :trunc],
            "length": torch.clamp(out["length"], 1, trunc),
        })
    return out


def frame_bce_loss(output):
    """losses.py:12-24."""
    loss = F.binary_cross_entropy(output["frame_sim"], output["label"], reduction="none")
    mask = generate_length_mask(output["length"]).to(loss.dtype)
    if mask.size(1) < loss.size(1):      # generate_length_mask uses max(length) columns
        loss = loss[:, :mask.size(1)]
    return (loss * mask).sum() / mask.sum()


TRAINABLE_KINDS = ("xavier", "bias", "gru", "emb", "bn_w", "bn_b")


def trainable_keys(vocab: int = VOCAB):
    return [k for k, _, kind in state_dict_spec(vocab) if kind in TRAINABLE_KINDS]


class AdamState:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) restated."""
    def __init__(self, keys):
        self.step = 0
        self.m = {k: None for k in keys}
        self.v = {k: None for k in keys}


def train_step(sd, batch, opt: AdamState, lr: float = 1e-3, max_grad_norm: float = 1.0,
               dropout: bool = False, dropout_masks=None, fast_gru: bool = False):
    """One iteration of run_strong.py:139-147: zero_grad, forward, loss, backward,
    clip_grad_norm_(max_grad_norm) (global L2, coef = max/(norm+1e-6) clamped to 1),
    Adam step.  Mutates ``sd`` in place.  Returns (loss, grads-before-clip, total_norm)."""
    keys = list(opt.m.keys())
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    out = runner_forward(sd, batch, training=True, dropout=dropout,
                         dropout_masks=dropout_masks, fast_gru=fast_gru)
    loss = frame_bce_loss(out)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_grad_norm / (total_norm + 1e-6), max=1.0)
    opt.step += 1
    b1, b2, eps = 0.9, 0.999, 1e-8
    bc1 = 1 - b1 ** opt.step
    bc2 = 1 - b2 ** opt.step
    with torch.no_grad():
        for k, p, g in zip(keys, params, grads):
            gc = g * coef
            if opt.m[k] is None:
                opt.m[k] = torch.zeros_like(p)
                opt.v[k] = torch.zeros_like(p)
            opt.m[k].mul_(b1).add_(gc, alpha=1 - b1)
            opt.v[k].mul_(b2).addcmul_(gc, gc, value=1 - b2)
            denom = (opt.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
            sd[k] = (p - (lr / bc1) * opt.m[k] / denom).detach()
    return loss.detach(), dict(zip(keys, grads)), total_norm


# --------------------------------------------------------------------------- weak (multi-phrase) path
# SURVEY.md §8f rank 1.  Reference sites restated here:
#   * MultiTextBiEncoder.forward ...... models/audio_text_model.py:147-229
#   * *_with_lens pooling ............. models/utils.py:33-95
#   * ClipBceLoss / ClipFrameBceLoss .. losses.py:38-43,186-210
#   * Runner.forward (weak) ........... python_scripts/training/run_weak_phrase.py:39-60
def sum_with_lens(features, lens):
    """models/utils.py:33-48."""
    mask = generate_length_mask(torch.as_tensor(lens), features.size(1))
    while mask.ndim < features.ndim:
        mask = mask.unsqueeze(-1)
    return (features * mask).sum(1)


def pool_with_lens(frame_sim, lens, pooling: str):
    """clip-level pooling of frame probabilities [B,T,n] over the valid frames (models/utils.py:51-95)."""
    lens = torch.as_tensor(lens)
    if pooling == "linear_softmax":
        return sum_with_lens(frame_sim ** 2, lens) / sum_with_lens(frame_sim, lens)
    if pooling == "mean":
        l = lens
        s = sum_with_lens(frame_sim, lens)
        while l.ndim < s.ndim:
            l = l.unsqueeze(1)
        return s / l
    if pooling == "max":
        mask = generate_length_mask(lens, frame_sim.size(1))
        f = frame_sim.clone()
        f[~mask] = float("-inf")
        return f.max(1)[0]
    if pooling == "exp_softmax":
        normed = frame_sim - frame_sim.max(1, keepdim=True)[0]
        e = torch.exp(normed)
        w = e / sum_with_lens(e, lens).unsqueeze(1)
        return sum_with_lens(w * frame_sim, lens)
    raise Exception(f"Unsupported pooling {pooling}")


def synth_weak_batch(batch: int, n_samples: int, n_phrases: int = 5, n_tokens: int = 6, seed: int = 0,
                     vocab: int = VOCAB, ragged: bool = True):
    """Seeded batch in the schema of the weak-phrase runner: text [B,n,N], text_len [B,n], weak label [B,n],
    strong label [B,T',n]."""
    base = synth_batch(batch, n_samples, n_tokens, seed, vocab, ragged)
    g = torch.Generator().manual_seed(seed + 1000)
    text = torch.randint(2, vocab, (batch, n_phrases, n_tokens), generator=g)
    text_len = torch.randint(1, n_tokens + 1, (batch, n_phrases), generator=g)
    for b in range(batch):
        for j in range(n_phrases):
            text[b, j, text_len[b, j]:] = 0
    t_out = (n_samples // HOP + 1) // 4
    base.update({
        "text": text, "text_len": text_len,
        "label": (torch.rand(batch, n_phrases, generator=g) > 0.5).float(),
        "strong_label": (torch.rand(batch, t_out, n_phrases, generator=g) > 0.5).float(),
    })
    base["weak_label"] = base["label"]
    return base


def multitext_forward(sd, input_dict, pooling: str = "linear_softmax", training=False, dropout=True,
                      dropout_masks=None, fast_gru=False):
    """models/audio_text_model.py:147-229 (no cross encoder / projections / upsample / safe_size)."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, None, fast_gru)
    audio_emb = a["embedding"]
    text = input_dict["text"]
    B, n = text.shape[0], text.shape[1]
    t = embedding_mean(sd, text.reshape(B * n, -1), torch.as_tensor(input_dict["text_len"]).reshape(B * n))
    audio_rep = audio_emb.unsqueeze(1).expand(-1, n, -1, -1).reshape(B * n, *audio_emb.shape[1:])
    frame_sim, _ = dot_product_match(audio_rep, t["seq_emb"])
    frame_sim = frame_sim.reshape(B, n, -1).transpose(1, 2)          # [B,T,n]
    clip_sim = pool_with_lens(frame_sim, a["length"], pooling)
    return {"frame_sim": frame_sim, "clip_sim": clip_sim, "length": a["length"]}


def clip_bce_loss(clip_sim, label):
    """losses.py:38-43."""
    return F.binary_cross_entropy(clip_sim, label)


def frame_bce_tensor(frame_sim, label, length):
    """losses.py:26-35 (FrameBceLoss.forward_tensor, 2-D or 3-D)."""
    loss = F.binary_cross_entropy(frame_sim, label, reduction="none")
    mask = generate_length_mask(length, loss.size(1)).to(loss.dtype)
    if loss.ndim == 3:
        mask = mask.unsqueeze(-1).expand(*loss.size())
    return (loss * mask).sum() / mask.sum()


def clip_frame_bce_loss(output, frame_weight: float):
    """losses.py:186-210 with the default keys."""
    return (1 - frame_weight) * clip_bce_loss(output["clip_sim"], output["weak_label"]) + \
        frame_weight * frame_bce_tensor(output["frame_sim"], output["strong_label"], output["length"])


# =====================================================================================================
# SURVEY.md §8f rank 3 — sentence-level alignment.  Reference sites restated here:
#   * align.DotProduct ................ models/align.py:7-31
#   * sim_pooling.Audio*Text* ......... models/sim_pooling.py:6-190
#   * AudioTextAlignByWord/ByPhrase ... models/audio_text_model.py:843-976
#   * MaxMarginRankingLoss ............ losses.py:226-264
def align_dot_product(audio, text, scaled: bool = False):
    """sim[i, j, t, n] for all (clip i, text j) pairs (models/align.py:14-31, l2norm=False)."""
    bs, n_seg, dim = audio.shape
    n_txt = text.shape[1]
    score = audio.reshape(-1, dim) @ text.reshape(-1, dim).t()
    if scaled:
        score = score / math.sqrt(dim)
    score = torch.sigmoid(score).clamp(1e-7, 1.0)
    return score.reshape(bs, n_seg, bs, n_txt).transpose(1, 2)


def _tokens_pool(x, lens, mode: str):
    """second reduction of sim_pooling: x [bs*bs, t_len] over n < lens."""
    lens = torch.as_tensor(lens)
    if mode == "mean":
        return sum_with_lens(x, lens) / lens
    if mode == "sum":
        return sum_with_lens(x, lens)
    if mode == "max":
        return pool_with_lens(x, lens, "max")
    if mode == "meansum":
        return sum_with_lens(x, lens) + sum_with_lens(x, lens) / lens
    raise Exception(mode)


def sim_pooling(sim, audio_len, text_len, audio_pool: str, text_pool: str):
    """models/sim_pooling.py: frames pooled with audio_len[i] (expanded over j), tokens with text_len[j]
    (repeated over i); returns [bs, bs]."""
    bs, a_len, t_len = sim.size(0), sim.size(2), sim.size(3)
    s = sim.reshape(bs * bs, a_len, t_len)
    al = torch.as_tensor(audio_len).unsqueeze(1).expand(bs, bs).reshape(-1)
    s = pool_with_lens(s, al, audio_pool)
    tl = torch.as_tensor(text_len).repeat(bs)
    return _tokens_pool(s, tl, text_pool).reshape(bs, bs)


def max_margin_ranking_loss(x, margin: float = 1.0, fix_norm: bool = True, lamda1: float = 1.0):
    """losses.py:235-264, written per element: both halves compare the positive x[i,i] with x[i,j] and with
    lamda1 * x[j,i]; fix_norm drops the i == j terms."""
    n = x.size(0)
    d = torch.diag(x).unsqueeze(1).expand(n, n)
    m1 = F.relu(margin - (d - x))
    m2 = F.relu(margin - (d - lamda1 * x.t()))
    if fix_norm:
        keep = 1.0 - torch.eye(n)
        return ((m1 * keep).sum() + (m2 * keep).sum()) / (2.0 * n * (n - 1))
    return (m1.sum() + m2.sum()) / (2.0 * n * n)


def synth_align_batch(batch: int, n_samples: int, max_phrases: int = 3, n_tokens: int = 6, seed: int = 0,
                      vocab: int = VOCAB):
    """Seeded batch in the schema of the sentence-level runners: word-level ``text`` [B, N] / ``text_len`` and
    phrase-level ``phrases`` [txt_num, N] / ``phrases_len`` / ``phrases_num`` (list, sums to txt_num)."""
    base = synth_batch(batch, n_samples, n_tokens, seed, vocab, True)
    g = torch.Generator().manual_seed(seed + 2000)
    num = torch.randint(1, max_phrases + 1, (batch,), generator=g)
    num[0] = max_phrases
    total = int(num.sum())
    phrases = torch.randint(2, vocab, (total, n_tokens), generator=g)
    plen = torch.randint(1, n_tokens + 1, (total,), generator=g)
    for r in range(total):
        phrases[r, plen[r]:] = 0
    base.update({"phrases": phrases, "phrases_len": plen, "phrases_num": num.tolist(), "text_key": "phrases"})
    return base


def align_forward(sd, input_dict, level: str = "phrase", audio_pool: str = "mean", text_pool: str = "mean",
                  scaled: bool = False, training=False, dropout=True, dropout_masks=None, fast_gru=False):
    """AudioTextAlignByWord.forward (models/audio_text_model.py:871-904) / AudioTextAlignByPhrase.forward
    (:937-976) without projections / cross encoder."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, None, fast_gru)
    if level == "word":
        t = embedding_mean(sd, input_dict["text"], torch.as_tensor(input_dict["text_len"]))
        text_emb, text_len = t["token_emb"], input_dict["text_len"]
    else:
        key = input_dict["text_key"]
        t = embedding_mean(sd, input_dict[key], torch.as_tensor(input_dict[f"{key}_len"]))
        num = [int(n) for n in input_dict[f"{key}_num"]]
        text_emb = torch.nn.utils.rnn.pad_sequence(torch.split(t["seq_emb"], num, dim=0), batch_first=True)
        text_len = num
    sim_matrix = align_dot_product(a["embedding"], text_emb, scaled)
    sim = sim_pooling(sim_matrix, a["length"], text_len, audio_pool, text_pool)
    return {"sim": sim, "sim_matrix": sim_matrix, "length": a["length"]}


# =====================================================================================================
# SURVEY.md §8f rank 2 — attention-type heads (BASELINE.json configs[3]).  Reference sites restated here:
#   * PositionalEncoding / SelfAttention ... models/text_encoder.py:128-146, 240-268
#   * nn.MultiheadAttention ................ torch 2.11 (packed in_proj, batch_first, key_padding_mask), used at
#                                            text_encoder.py:258-266 and match.py:66-73,82
#   * match.CrossAttention ................. models/match.py:63-88
#   * DotProduct(text_level="token") ....... models/match.py:43-60
#   * Seq2SeqAttention / CrossGating / CrossAttentionGating ... models/cross_encoder.py:5-79
#   * BiEncoder.forward with a cross encoder ... models/audio_text_model.py:58-98
HEADS = 8


def positional_table(max_len: int = 100, d_model: int = EMBED) -> torch.Tensor:
    """models/text_encoder.py:132-140 -> [1, max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def attn_state_spec(E: int = EMBED):
    te, mf, ce = "text_encoder.", "match_fn.", "cross_encoder."
    return {
        "selfattn": [(te + "cls_token", (1, 1, E), 0.5), (te + "mha.in_proj_weight", (3 * E, E), None),
                     (te + "mha.in_proj_bias", (3 * E,), 0.1), (te + "mha.out_proj.weight", (E, E), None),
                     (te + "mha.out_proj.bias", (E,), 0.1)],
        "crossattn": [(mf + "attn.in_proj_weight", (3 * E, E), None), (mf + "attn.in_proj_bias", (3 * E,), 0.1),
                      (mf + "attn.out_proj.weight", (E, E), None), (mf + "attn.out_proj.bias", (E,), 0.1),
                      (mf + "norm.weight", (E,), "ln_w"), (mf + "norm.bias", (E,), 0.2),
                      (mf + "linear.weight", (1, E), 0.3), (mf + "linear.bias", (1,), 0.1)],
        "gating": [(ce + "attn.h2attn.weight", (E, 2 * E), None), (ce + "attn.h2attn.bias", (E,), 0.1),
                   (ce + "attn.v", (E,), "randn"), (ce + "gating.fc_u.weight", (E, E), None),
                   (ce + "gating.fc_u.bias", (E,), 0.1), (ce + "gating.fc_s.weight", (E, E), None),
                   (ce + "gating.fc_s.bias", (E,), 0.1)],
    }


def synth_attn_state(seed: int, parts=("selfattn", "crossattn", "gating"), gain: float = 1.0):
    """Deterministic weights for the attention-type heads (Xavier-uniform matrices scaled by ``gain``, small
    non-zero biases so no term is vacuous).  ``text_encoder.pe.pe`` is the sinusoidal buffer."""
    sd = {}
    spec = attn_state_spec()
    idx = 0
    for part in ("selfattn", "crossattn", "gating"):
        for key, shape, kind in spec[part]:
            idx += 1
            if part not in parts:
                continue
            g = torch.Generator().manual_seed(seed * 7919 + idx)
            if kind is None:
                a = gain * math.sqrt(6.0 / (shape[0] + shape[1]))
                t = (torch.rand(shape, generator=g) * 2 - 1) * a
            elif kind == "ln_w":
                t = 0.75 + 0.5 * torch.rand(shape, generator=g)
            elif kind == "randn":
                t = torch.randn(shape, generator=g)
            else:
                t = (torch.rand(shape, generator=g) * 2 - 1) * kind
            sd[key] = t
    if "selfattn" in parts:
        sd["text_encoder.pe.pe"] = positional_table()
    return sd


def attn_case_state(variant: str, seed: int, attn_seed: int, sharpen: float = 20.0):
    """Weights of one parity case: audio encoder + embedding from synth_state_dict, heads from synth_attn_state.
    For the "gating" variant the text encoder's out_proj is scaled by 25 so that the scaled dot-product logits
    span about +-3 instead of +-0.08 (a vacuous comparison on frame_sim otherwise)."""
    sd = synth_state_dict(seed=seed, sharpen=sharpen, perturb_bn=True)
    parts = ("selfattn", "crossattn") if variant == "crossattn" else ("selfattn", "gating")
    sd.update(synth_attn_state(attn_seed, parts))
    if variant == "gating":
        sd["text_encoder.mha.out_proj.weight"] = sd["text_encoder.mha.out_proj.weight"] * 25.0
        sd["text_encoder.mha.out_proj.bias"] = sd["text_encoder.mha.out_proj.bias"] * 25.0
    return sd


def multi_head_attention(xq, xkv, in_w, in_b, out_w, out_b, heads: int, key_len):
    """nn.MultiheadAttention forward (eval / dropout off): packed in_proj, q scaled by 1/sqrt(dh), keys
    n >= key_len[b] masked to -inf, softmax, heads concatenated, out_proj."""
    B, Lq, E = xq.shape
    Lk = xkv.shape[1]
    dh = E // heads
    q = xq @ in_w[:E].t() + in_b[:E]
    k = xkv @ in_w[E:2 * E].t() + in_b[E:2 * E]
    v = xkv @ in_w[2 * E:].t() + in_b[2 * E:]
    q = q.view(B, Lq, heads, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(B, Lk, heads, dh).transpose(1, 2)
    v = v.view(B, Lk, heads, dh).transpose(1, 2)
    score = q @ k.transpose(-1, -2)                                    # [B,h,Lq,Lk]
    mask = ~generate_length_mask(torch.as_tensor(key_len), Lk)         # True = padded
    score = score.masked_fill(mask[:, None, None, :], float("-inf"))
    ctx = torch.softmax(score, dim=-1) @ v
    ctx = ctx.transpose(1, 2).reshape(B, Lq, E)
    return ctx @ out_w.t() + out_b


def self_attention_text_encoder(sd, text, text_len, heads: int = HEADS):
    """models/text_encoder.py:260-268 (dropout off)."""
    p = "text_encoder."
    x = F.embedding(text.long(), sd[p + "embedding.core.weight"])
    x = torch.cat((sd[p + "cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    x = x + sd[p + "pe.pe"][:, :x.size(1)]
    lens = torch.as_tensor(text_len) + 1
    x = multi_head_attention(x, x, sd[p + "mha.in_proj_weight"], sd[p + "mha.in_proj_bias"],
                             sd[p + "mha.out_proj.weight"], sd[p + "mha.out_proj.bias"], heads, lens)
    return {"token_emb": x[:, 1:], "seq_emb": x[:, 0]}


def cross_attention_match(sd, audio, token_emb, text_len, heads: int = HEADS):
    """models/match.py:76-88 (dropout off)."""
    p = "match_fn."
    out = multi_head_attention(audio, token_emb, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
                               sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], heads, text_len)
    out = F.layer_norm(audio + out, (audio.size(-1),), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    out = out @ sd[p + "linear.weight"].t() + sd[p + "linear.bias"]
    return torch.sigmoid(out).squeeze(-1)


def cross_attention_gating(sd, audio, token_emb, audio_len, text_len):
    """models/cross_encoder.py:12-42 (additive attention, written without the [B, T*N, 2E] expansion: h2attn is
    linear in cat(q, kv)), :52-57 (gating), :67-79."""
    p = "cross_encoder."
    B, T, E = audio.shape
    N = token_emb.shape[1]
    W, b, v = sd[p + "attn.h2attn.weight"], sd[p + "attn.h2attn.bias"], sd[p + "attn.v"]
    hq = audio @ W[:, :E].t()
    hk = token_emb @ W[:, E:].t() + b
    score = (torch.tanh(hq.unsqueeze(2) + hk.unsqueeze(1)) * v).sum(-1)            # [B,T,N]
    m1 = generate_length_mask(torch.as_tensor(audio_len), T).unsqueeze(-1)
    m2 = generate_length_mask(torch.as_tensor(text_len), N).unsqueeze(1)
    score = score.masked_fill(~m1, -1e10).masked_fill(~m2, -1e10)
    text_att = torch.softmax(score, dim=-1) @ token_emb                            # [B,T,E]
    g_u = torch.sigmoid(audio @ sd[p + "gating.fc_u.weight"].t() + sd[p + "gating.fc_u.bias"])
    s_out = text_att * g_u
    g_s = torch.sigmoid(text_att @ sd[p + "gating.fc_s.weight"].t() + sd[p + "gating.fc_s.bias"])
    u_out = audio * g_s
    return u_out, s_out


def attn_biencoder_forward(sd, input_dict, variant: str, training=False, dropout=True, dropout_masks=None,
                           fast_gru=False):
    """BiEncoder.forward (models/audio_text_model.py:58-98) for the attention configurations:
    variant "crossattn": SelfAttention text encoder + match.CrossAttention;
    variant "gating":    SelfAttention text encoder + CrossAttentionGating + DotProduct(text_level="token").
    Dropout inside the attention heads is off (parity runs patch F.dropout / use eval for those modules)."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, None, fast_gru)
    t = self_attention_text_encoder(sd, input_dict["text"], input_dict["text_len"])
    audio = a["embedding"]
    if variant == "crossattn":
        sim = cross_attention_match(sd, audio, t["token_emb"], input_dict["text_len"])
    else:
        u, s = cross_attention_gating(sd, audio, t["token_emb"], a["length"], input_dict["text_len"])
        sim = torch.sigmoid((u * s).sum(-1) / math.sqrt(u.size(-1))).clamp(1e-7, 1.0)     # match.py:43-60, "token"
    return {"frame_sim": sim, "length": a["length"], "token_emb": t["token_emb"], "seq_emb": t["seq_emb"]}


# =====================================================================================================
# BASELINE.json configs[4] — CLAP text tower behind the Hugging Face surface.  The arithmetic of the tower lives in
# the `transformers` library (ClapTextModel / ClapProjectionLayer, version pinned by the image: the reference's
# requirements.txt does not list it); the oracle for it IS that library code run on CPU in fp32, wired exactly as
# the reference's LaionClapEncoder.forward does (models/hf_modeling_grounding.py:192-199).
def clap_case_modules(tower_seed: int):
    """(text tower, projection, extra state) of one parity case: full-size randomly initialised ClapTextModel +
    ClapProjectionLayer under ``tower_seed`` and Xavier audio_proj / 100x-sharpened text_proj (seq_emb is unit-norm,
    so the raw logits would sit in +-0.05)."""
    from transformers import ClapTextConfig
    from transformers.models.clap import modeling_clap as mc
    torch.manual_seed(tower_seed)
    cfg = ClapTextConfig()
    tower = mc.ClapTextModel(cfg).eval()
    proj = mc.ClapProjectionLayer(cfg).eval()
    g = torch.Generator().manual_seed(tower_seed + 1)
    a = math.sqrt(6.0 / (EMBED + EMBED))
    extra = {
        "audio_proj.weight": (torch.rand(EMBED, EMBED, generator=g) * 2 - 1) * a,
        "audio_proj.bias": (torch.rand(EMBED, generator=g) * 2 - 1) * 0.1,
        "text_proj.weight": (torch.rand(EMBED, EMBED, generator=g) * 2 - 1) * a * 100.0,
        "text_proj.bias": (torch.rand(EMBED, generator=g) * 2 - 1) * 0.1,
    }
    return tower, proj, extra


def synth_clap_batch(batch: int, n_samples: int, n_tokens: int, seed: int):
    """waveforms as synth_batch; RoBERTa-style ids: <s>=0 ... </s>=2, right padded with pad id 1."""
    base = synth_batch(batch, n_samples, 8, seed, ragged=True)
    g = torch.Generator().manual_seed(seed + 3000)
    lens = torch.randint(3, n_tokens + 1, (batch,), generator=g)
    lens[0] = n_tokens
    ids = torch.full((batch, n_tokens), 1, dtype=torch.long)
    mask = torch.zeros(batch, n_tokens, dtype=torch.long)
    for b in range(batch):
        n = int(lens[b])
        ids[b, 0] = 0
        ids[b, 1:n - 1] = torch.randint(4, 50000, (n - 2,), generator=g)
        ids[b, n - 1] = 2
        mask[b, :n] = 1
    return {"waveform": base["waveform"], "waveform_len": base["waveform_len"], "input_ids": ids,
            "attention_mask": mask}


def clap_text_encoder(tower, proj, input_ids, attention_mask):
    """LaionClapEncoder.forward (models/hf_modeling_grounding.py:192-199) on the library modules."""
    out = tower(input_ids=input_ids.long(), attention_mask=attention_mask.long())
    token_emb = proj(out.last_hidden_state)
    seq_emb = F.normalize(proj(out.pooler_output), dim=-1)
    return {"seq_emb": seq_emb, "token_emb": token_emb}


# =====================================================================================================
# SURVEY.md §8f rank 4(iii) — post-processing of frame probabilities.  Reference sites restated here (numpy):
#   * find_contiguous_regions ... utils/eval_util.py:18-44
#   * median_filter (+ binarize) . utils/eval_util.py:47-63 (sklearn binarize: x > th; scipy.ndimage.median_filter,
#                                  size (1, w), mode="reflect": window t - w//2 .. t - w//2 + w - 1, rank w // 2)
#   * connect_clusters / connect_  utils/eval_util.py:74-116
#   * the per-sample, per-threshold loop ... python_scripts/training/run_strong.py:222-247
def contiguous_regions(activity):
    import numpy as np
    a = np.asarray(activity).astype(bool)
    change = np.logical_xor(a[1:], a[:-1]).nonzero()[0] + 1
    if a[0]:
        change = np.r_[0, change]
    if a[-1]:
        change = np.r_[change, a.size]
    return change.reshape((-1, 2))


def median_filter_binary(x, window: int, threshold: float):
    import numpy as np
    x = np.asarray(x)
    b = (x.astype(np.float64) > threshold).astype(np.int64)
    T = b.size
    out = np.zeros(T, dtype=np.int64)
    half = window // 2
    for t in range(T):
        vals = []
        for i in range(window):
            idx = t - half + i
            if idx < 0:
                idx = -idx - 1
            if idx >= T:
                idx = 2 * T - idx - 1
            idx = min(max(idx, 0), T - 1)
            vals.append(b[idx])
        out[t] = sorted(vals)[window // 2]
    return out


def connect_regions(pairs, n: int):
    pairs = [tuple(int(v) for v in p) for p in pairs]
    if not pairs:
        return []
    merged = []
    start, end = pairs[0]
    for s, e in pairs[1:]:
        if s - end <= n:
            end = e
        else:
            merged.append((start, end))
            start, end = s, e
    merged.append((start, end))
    return merged


def frame_regions(frame_sim_row, threshold: float, window: int, n_connect: int):
    """(onset, offset) frame pairs of one sample at one threshold (run_strong.py:231-241)."""
    return connect_regions(contiguous_regions(median_filter_binary(frame_sim_row, window, threshold)), n_connect)


# =====================================================================================================
# Normalised match heads.  Reference sites: DotProduct(l2norm=True) models/match.py:43-60 (F.normalize both sides =
# cosine similarity), ExpNegL2 models/match.py:10-33.
def dot_product_match_l2norm(audio_emb, seq_emb, scale: bool = True):
    a = F.normalize(audio_emb, dim=-1)
    s = F.normalize(seq_emb, dim=-1).unsqueeze(1)
    score = (a * s).sum(-1)
    if scale:
        score = score / math.sqrt(audio_emb.size(-1))
    return torch.sigmoid(score).clamp(1e-7, 1.0)


def exp_neg_l2_match(audio_emb, seq_emb, l2norm: bool = True):
    a, s = audio_emb, seq_emb
    if l2norm:
        a, s = F.normalize(a, dim=-1), F.normalize(s, dim=-1)
    return torch.exp(-torch.norm(a - s.unsqueeze(1), dim=-1))


# =====================================================================================================
# Options of the same model classes reachable from the run_strong.py / run_weak_phrase.py config surface:
#   * BiEncoder / MultiTextBiEncoder(add_proj=True) ...... models/audio_text_model.py:35-46, 79-88, 150-151, 180-185
#   * MultiTextBiEncoder(cross_encoder=...) ............... models/audio_text_model.py:166-178
#   * upsample=True (F.interpolate linear) ............... models/audio_text_model.py:90-97, 216-223
#   * EmbeddingAgg(aggregation="attention") .............. models/text_encoder.py:46-58, 84-85
def proj_state(seed: int, E: int = EMBED, gain: float = 1.0):
    """audio_proj / text_proj = nn.Linear(E, shared_dim = E): Xavier-uniform matrices (x gain), small biases."""
    sd = {}
    for idx, key in enumerate(("audio_proj", "text_proj")):
        g = torch.Generator().manual_seed(seed * 4099 + idx)
        a = gain * math.sqrt(6.0 / (2 * E))
        sd[key + ".weight"] = (torch.rand(E, E, generator=g) * 2 - 1) * a
        sd[key + ".bias"] = (torch.rand(E, generator=g) * 2 - 1) * 0.1
    return sd


def attn_pool_state(seed: int, E: int = EMBED, gain: float = 1.0):
    """text_encoder.attn.fc = nn.Linear(E, 1) of AttentionPooling."""
    g = torch.Generator().manual_seed(seed * 5003 + 1)
    return {"text_encoder.attn.fc.weight": (torch.rand(1, E, generator=g) * 2 - 1) * gain * math.sqrt(6.0 / E),
            "text_encoder.attn.fc.bias": (torch.rand(1, generator=g) * 2 - 1) * 0.1}


def attention_pooling(sd, emb, lens):
    """models/text_encoder.py:51-58: masked (-1e10) softmax over the tokens of fc(x), weighted sum of x."""
    w, b = sd["text_encoder.attn.fc.weight"], sd["text_encoder.attn.fc.bias"]
    score = (emb * w.view(1, 1, -1)).sum(-1) + b
    mask = generate_length_mask(torch.as_tensor(lens), emb.size(1))
    score = score.masked_fill(~mask, -1e10)
    weight = torch.softmax(score, dim=1)
    return (emb * weight.unsqueeze(-1)).sum(1)


def linear_upsample(x, size: int):
    """F.interpolate(size=size, mode="linear", align_corners=False) along dim 1 of x [B,T] or [B,T,n], written out
    (ATen upsample_linear1d): scale = T/size, src = scale*(dst + 0.5) - 0.5 clamped at 0, neighbours
    i0 = min(floor(src), T-1), i1 = min(i0+1, T-1), weight lam = src - i0."""
    T = x.shape[1]
    dst = torch.arange(size, dtype=torch.float32)
    scale = torch.tensor(float(T)) / torch.tensor(float(size))          # float32 division, as ATen
    src = (scale * (dst + 0.5) - 0.5).clamp_min(0.0)
    i0 = torch.clamp(src.floor().long(), max=T - 1)
    i1 = torch.clamp(i0 + 1, max=T - 1)
    lam = (src - i0.to(src.dtype)).clamp(0.0, 1.0)
    lam = lam.view([1, -1] + [1] * (x.ndim - 2))
    return (1.0 - lam) * x.index_select(1, i0) + lam * x.index_select(1, i1)


def biencoder_variant_forward(sd, input_dict, training=False, dropout=True, dropout_masks=None, fast_gru=False,
                              add_proj=False, upsample=False, aggregation="mean"):
    """BiEncoder.forward (models/audio_text_model.py:58-98) with its options: projections after the encoders,
    attention aggregation of the word embeddings, x4 linear upsampling of frame_sim."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, None, fast_gru)
    t = embedding_mean(sd, input_dict["text"], input_dict["text_len"])
    if aggregation == "attention":
        t["seq_emb"] = attention_pooling(sd, t["token_emb"], input_dict["text_len"])
    audio, seq = a["embedding"], t["seq_emb"]
    if add_proj:
        audio = audio @ sd["audio_proj.weight"].t() + sd["audio_proj.bias"]
        seq = seq @ sd["text_proj.weight"].t() + sd["text_proj.bias"]
    frame_sim, _ = dot_product_match(audio, seq)
    length = a["length"]
    if upsample:
        frame_sim = linear_upsample(frame_sim, frame_sim.size(1) * 4)
        length = length * 4
    return {"frame_sim": frame_sim, "length": length}


def multitext_variant_forward(sd, input_dict, pooling="linear_softmax", training=False, dropout=True,
                              dropout_masks=None, fast_gru=False, add_proj=False, cross=False, upsample=False):
    """MultiTextBiEncoder.forward (models/audio_text_model.py:147-229) with its options.  Order: audio_proj on the
    audio embedding first (:150-151), text encoder on [B*n, N], expansion to one (clip, phrase) pair per row
    (:166-170), cross encoder (:177-179), text_proj (:180-185), match, pooling, upsampling (:216-223)."""
    a = cnn8rnn_forward(sd, input_dict["waveform"], input_dict["waveform_len"], training,
                        dropout_masks, dropout, None, fast_gru)
    audio = a["embedding"]
    if add_proj:
        audio = audio @ sd["audio_proj.weight"].t() + sd["audio_proj.bias"]
    text = input_dict["text"]
    B, n = text.shape[0], text.shape[1]
    text_len = torch.as_tensor(input_dict["text_len"]).reshape(B * n)
    t = embedding_mean(sd, text.reshape(B * n, -1), text_len)
    audio_rep = audio.unsqueeze(1).expand(-1, n, -1, -1).reshape(B * n, *audio.shape[1:])
    audio_len = torch.as_tensor(a["length"]).repeat_interleave(n)
    if cross:
        u, s = cross_attention_gating(sd, audio_rep, t["token_emb"], audio_len, text_len)
        if add_proj:
            s = s @ sd["text_proj.weight"].t() + sd["text_proj.bias"]
        frame_sim = torch.sigmoid((u * s).sum(-1) / math.sqrt(u.size(-1))).clamp(1e-7, 1.0)
    else:
        seq = t["seq_emb"]
        if add_proj:
            seq = seq @ sd["text_proj.weight"].t() + sd["text_proj.bias"]
        frame_sim, _ = dot_product_match(audio_rep, seq)
    frame_sim = frame_sim.reshape(B, n, -1).transpose(1, 2)
    length = a["length"]
    clip_sim = pool_with_lens(frame_sim, length, pooling)
    if upsample:
        # reference quirk: the target size is frame_sim.size(-1) * ratio on the [B,T,n] tensor = n * 4 frames (:217-219)
        frame_sim = linear_upsample(frame_sim, frame_sim.size(-1) * 4)
        length = length * 4
    return {"frame_sim": frame_sim, "clip_sim": clip_sim, "length": length}


VARIANT_CASE = dict(batch=3, n_samples=32000, n_phrases=4, n_tokens=6, seed=8, data_seed=9, head_seed=17)


def variant_state(variant: str):
    """Weights of one option-parity case (tests/golden/variants_b3_1s.npz)."""
    c = VARIANT_CASE
    if variant in ("multi_proj", "bi_proj_upsample"):
        sd = synth_state_dict(seed=c["seed"], sharpen=100.0, perturb_bn=True)
        sd.update(proj_state(c["head_seed"]))
    elif variant == "multi_gating_proj":
        sd = synth_state_dict(seed=c["seed"], sharpen=20.0, perturb_bn=True)
        sd.update(proj_state(c["head_seed"], gain=4.0))
        sd.update(synth_attn_state(c["head_seed"], ("gating",)))
    elif variant == "bi_attnagg":
        sd = synth_state_dict(seed=c["seed"], sharpen=100.0, perturb_bn=True)
        sd.update(attn_pool_state(c["head_seed"], gain=0.05))
    elif variant == "multi_upsample":
        sd = synth_state_dict(seed=c["seed"], sharpen=100.0, perturb_bn=True)
    else:
        raise KeyError(variant)
    return sd


def variant_forward(variant: str, sd, batch, **kw):
    if variant == "multi_proj":
        return multitext_variant_forward(sd, batch, add_proj=True, **kw)
    if variant == "multi_gating_proj":
        return multitext_variant_forward(sd, batch, add_proj=True, cross=True, **kw)
    if variant == "multi_upsample":
        return multitext_variant_forward(sd, batch, upsample=True, **kw)
    if variant == "bi_proj_upsample":
        return biencoder_variant_forward(sd, batch, add_proj=True, upsample=True, **kw)
    if variant == "bi_attnagg":
        return biencoder_variant_forward(sd, batch, aggregation="attention", **kw)
    raise KeyError(variant)


def variant_batch(variant: str):
    c = VARIANT_CASE
    if variant.startswith("multi"):
        return synth_weak_batch(c["batch"], c["n_samples"], c["n_phrases"], c["n_tokens"], seed=c["data_seed"])
    b = synth_batch(c["batch"], c["n_samples"], c["n_tokens"], seed=c["data_seed"], ragged=True)
    if variant == "bi_proj_upsample":        # labels at the upsampled resolution
        g = torch.Generator().manual_seed(c["data_seed"] + 77)
        t_out = (c["n_samples"] // HOP + 1) // 4
        b["label"] = (torch.rand(c["batch"], 4 * t_out + 1, generator=g) > 0.5).float()
    return b
