"""oracle/stubs: LPIPS needs pretrained AlexNet weights (no network here): the stand-in returns 0 and announces it."""
import sys

import torch


class LearnedPerceptualImagePatchSimilarity:
    def __init__(self, *a, **k):
        print("[oracle/stubs] LPIPS unavailable (no pretrained weights): reporting 0", file=sys.stderr)

    def to(self, *a, **k):
        return self

    def __call__(self, a, b):
        return torch.zeros((), device=a.device)
