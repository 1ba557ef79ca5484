"""Host-side construction of the frontend constant buffers (Hann window, slaney mel
filterbank) — what torchaudio.transforms.MelSpectrogram computes at construction time for the
reference (models/audio_encoder.py:113-123: n_fft=win=1024, hop=320, f_min=50, f_max=14000,
n_mels=64, norm="slaney", mel_scale="slaney").  Pure setup code; the per-step arithmetic is in
csrc/frontend.cu."""
import math

import torch


def hann_window(n: int = 1024) -> torch.Tensor:
    """Periodic Hann window, as torch.hann_window(n)."""
    k = torch.arange(n, dtype=torch.float64, device="cpu")      # "cpu": also under a meta-device init context
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(torch.float32)


def _hz_to_mel(f: float) -> float:
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f / f_sp


def _mel_to_hz(m: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    hz = f_sp * m
    is_log = m >= min_log_mel
    hz[is_log] = 1000.0 * torch.exp(logstep * (m[is_log] - min_log_mel))
    return hz


def slaney_mel_fbanks(n_freqs: int = 513, f_min: float = 50.0, f_max: float = 14000.0,
                      n_mels: int = 64, sample_rate: int = 32000) -> torch.Tensor:
    """[n_freqs, n_mels] triangular filters, slaney mel scale and slaney (area) normalisation."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs, device="cpu")
    m_pts = torch.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2, device="cpu")
    f_pts = _mel_to_hz(m_pts)
    width = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
    rising = -slopes[:, :-2] / width[:-1]
    falling = slopes[:, 2:] / width[1:]
    fb = torch.clamp(torch.minimum(rising, falling), min=0.0)
    area_norm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * area_norm.unsqueeze(0)
