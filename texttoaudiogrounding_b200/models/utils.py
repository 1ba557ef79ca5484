"""Host-side helpers mirroring reference models/utils.py (init_weights 5-19,
generate_length_mask 22-30)."""
import torch
import torch.nn as nn


def init_weights(m):
    if isinstance(m, (nn.Conv2d, nn.Conv1d)):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
        nn.init.kaiming_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Embedding):
        nn.init.kaiming_uniform_(m.weight)


def generate_length_mask(lens, max_length=None):
    lens = torch.as_tensor(lens)
    if max_length is None:
        max_length = int(lens.max().item())
    idxs = torch.arange(max_length, device=lens.device).unsqueeze(0)
    return idxs < lens.view(-1, 1)


def lens_to_device(lens, device) -> torch.Tensor:
    """list / numpy / int or float tensor (the reference's Runner.forward hands text_len over as
    float32 on device, run_strong.py:94-99) -> int64 tensor on ``device``."""
    t = torch.as_tensor(lens)
    return t.to(device=device, dtype=torch.long)
