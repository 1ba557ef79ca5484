"""Similarity-matrix pooling — host-side mirror of reference models/sim_pooling.py:6-204.

Each class names a (frame pooling, token pooling) pair applied to sim[i, j, t, n]; with a
:class:`~.align.PairwiseSim` input both reductions run fused with the dot products in csrc/align.cu (the 4-D
matrix is never written).  ``MultiText*`` pool an ordinary [B, T, n] tensor over its frames."""
from __future__ import annotations

import torch.nn as nn

from . import utils
from .align import PairwiseSim


class _AudioTextPooling(nn.Module):
    audio_pool = "mean"
    text_pool = "mean"

    def forward(self, input):
        sim = input["sim"]
        if not isinstance(sim, PairwiseSim):
            raise NotImplementedError("sim_pooling (B200) pools the deferred similarity of models.align.DotProduct; "
                                      "a materialised 4-D tensor has no fused path")
        return sim.pool(input["audio_len"], input["text_len"], self.audio_pool, self.text_pool)


class AudioMeanTextMean(_AudioTextPooling):
    audio_pool, text_pool = "mean", "mean"


class AudioMeanTextSum(_AudioTextPooling):
    audio_pool, text_pool = "mean", "sum"


class AudioMaxTextMean(_AudioTextPooling):
    audio_pool, text_pool = "max", "mean"


class AudioMaxTextMax(_AudioTextPooling):
    audio_pool, text_pool = "max", "max"


class AudioMaxTextSum(_AudioTextPooling):
    audio_pool, text_pool = "max", "sum"


class AudioMaxTextMeanSum(_AudioTextPooling):
    audio_pool, text_pool = "max", "meansum"


class AudioLinearSoftTextMean(_AudioTextPooling):
    audio_pool, text_pool = "linear_softmax", "mean"


class AudioLinearSoftTextSum(_AudioTextPooling):
    audio_pool, text_pool = "linear_softmax", "sum"


class AudioExpSoftTextMean(_AudioTextPooling):
    audio_pool, text_pool = "exp_softmax", "mean"


class AudioExpSoftTextSum(_AudioTextPooling):
    audio_pool, text_pool = "exp_softmax", "sum"


class MultiTextLinearSoft(nn.Module):
    def forward(self, input):
        # reference: sim [bs, n_txt, n_seg] -> transpose(1, 2) -> pooled over the frames
        return utils.linear_softmax_with_lens(input["sim"].transpose(1, 2), input["audio_len"])


class MultiTextMax(nn.Module):
    def forward(self, input):
        return utils.max_with_lens(input["sim"].transpose(1, 2), input["audio_len"])
