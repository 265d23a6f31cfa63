"""sequence_mask with the reference's signature (blvm/utils/operations.py:90-119).  The kernels never materialise
this mask (they take per-utterance lengths); it exists for callers that want the tensor (e.g. VRNN returns it)."""
import math
from typing import Union

import torch

__all__ = ["sequence_mask", "level_lengths"]


def sequence_mask(seq_lens: Union[list, torch.Tensor], stride: int = 1, max_len: int = None,
                  dtype: torch.dtype = torch.bool, device: torch.device = None):
    if isinstance(seq_lens, torch.Tensor):
        device = seq_lens.device if device is None else device
        if device != seq_lens.device:
            seq_lens = seq_lens.to(device)
    else:
        seq_lens = torch.tensor(seq_lens, device=device, dtype=int)
    T = max_len or math.ceil(seq_lens.max() / stride)
    return (torch.arange(T, device=device).unsqueeze(0) < seq_lens.unsqueeze(1)).to(dtype)


def level_lengths(x_sl: torch.Tensor, stride: int) -> torch.Tensor:
    """Valid latent steps per utterance for a latent layer of temporal stride `stride`: ceil(x_sl / stride).
    Equals `seq_mask[:, ::stride].sum(1)` (vrnn.py:271, stcn.py:284) and the level masks of clockwork_vae.py:237-238."""
    return (x_sl + (stride - 1)) // stride
