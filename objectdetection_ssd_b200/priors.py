"""Prior (default box) tables, generated on the host.

``create_priors_ssd300`` must be bit-identical to the reference (Util.py:105-137): the centre
and size of every prior are computed in Python float64, rounded to float32 once, and clamped to
[0, 1] in cx,cy,w,h form (so the corner form can leave the unit square).  It is a one-off
8732 x 4 table; computing it on the GPU in fp32 would change low bits and with them match indices.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import torch


@dataclass(frozen=True)
class PriorSpec:
    grids: List[int]
    scales: List[float]
    ratios: List[List[float]] = field(default_factory=list)
    last_extra_scale: float = 1.0      # Util.py:131-132: the last level has no s_{k+1}

    @property
    def num_priors(self) -> int:
        return sum(g * g * (len(r) + 1) for g, r in zip(self.grids, self.ratios))


_R3 = [1., 2., 0.5]
_R5 = [1., 2., 3., 0.5, .333]          # .333, not 1/3 (Util.py:114-116)

SSD300_SPEC = PriorSpec(grids=[38, 19, 10, 5, 3, 1],
                        scales=[0.1, 0.2, 0.375, 0.55, 0.725, 0.9],
                        ratios=[_R3, _R5, _R5, _R5, _R3, _R3])
# SSD512-style table of the stress configuration (BASELINE.json configs[4]): 24 564 priors
SSD512_SPEC = PriorSpec(grids=[64, 32, 16, 8, 4, 2, 1],
                        scales=[0.07, 0.15, 0.3, 0.45, 0.6, 0.75, 0.9],
                        ratios=[_R3, _R5, _R5, _R5, _R5, _R3, _R3])


def make_priors(spec: PriorSpec = SSD300_SPEC) -> torch.Tensor:
    """[P,4] float32 cx,cy,w,h; order level -> row -> column -> ratio, the extra square prior of
    scale sqrt(s_k s_{k+1}) sits right after ratio 1 (Util.py:120-134)."""
    table = []
    levels = len(spec.grids)
    for k in range(levels):
        g = spec.grids[k]
        s = spec.scales[k]
        extra = math.sqrt(s * spec.scales[k + 1]) if k + 1 < levels else spec.last_extra_scale
        shapes = []
        for a in spec.ratios[k]:
            shapes.append((s * math.sqrt(a), s / math.sqrt(a)))
            if a == 1.:
                shapes.append((extra, extra))
        centres = [(c + 0.5) / float(g) for c in range(g)]
        for cy in centres:
            for cx in centres:
                table.extend([cx, cy, w, h] for (w, h) in shapes)
    out = torch.tensor(table, dtype=torch.float64).to(torch.float32)
    return out.clamp_(0, 1)


def cxcywh_to_xyxy_host(p: torch.Tensor) -> torch.Tensor:
    """Corner form of a (CPU) prior table with the reference's fp32 ops (Util.py:93-96)."""
    half = p[:, 2:] / 2.
    return torch.cat((p[:, :2] - half, p[:, :2] + half), dim=1).contiguous()
