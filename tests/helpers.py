"""Shared input builders for the parity tests (seeded, identical bits for oracle and kernels)."""
import numpy as np
import torch

from objectdetection_ssd_b200 import synth
from oracle import ssd_oracle as O


def priors(kind="ssd300"):
    if kind == "ssd300":
        return O.make_priors()
    if kind == "ssd512":
        return O.make_priors(**O.SSD512)
    raise ValueError(kind)


def train_inputs(seed, B, P, min_gt=1, max_gt=10, C=21):
    gb, gc = synth.make_gt(seed, B, min_gt, max_gt)
    loc, conf = synth.make_head(seed, B, P, C)
    tb = [torch.from_numpy(b) for b in gb]
    tc = [torch.from_numpy(c) for c in gc]
    return torch.from_numpy(loc), torch.from_numpy(conf), tb, tc


def detect_inputs(seed, B, P, bg_bias=6.0, C=21):
    loc, conf = synth.make_head(seed, B, P, C, loc_scale=0.5, bg_bias=bg_bias)
    return torch.from_numpy(loc), torch.from_numpy(conf)


def unpack_mask(mask_i32: torch.Tensor, P: int) -> torch.Tensor:
    """uint32 bit mask [B, ceil(P/32)] (stored as int32) -> bool [B,P]."""
    m = mask_i32.cpu().numpy().view(np.uint32)
    bits = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(bool)
    return torch.from_numpy(bits.reshape(m.shape[0], -1)[:, :P])
