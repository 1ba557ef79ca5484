"""B200-native (sm_100a) implementation of the TextToAudioGrounding cnn8rnn-w2vmean hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every stage of the
forward/backward step runs in hand-written CUDA kernels behind the C ABI declared in
``include/tag_b200.h`` (``lib/libtag_b200.so``).  The module classes under
``texttoaudiogrounding_b200.models`` / ``.losses`` mirror the reference's plugin surface
(dotted-path YAML instantiation, dict-in/dict-out contracts, identical state-dict keys).
"""
__version__ = "0.1.0"
