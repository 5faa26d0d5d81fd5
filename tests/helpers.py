"""Shared helpers for the parity tests: golden fixture loading and replay."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RACE_CASES = ['predef', 'iid9', 'floatw', 'loops', 'p1_crash', 'p4_short', 'agents']


def load_case(name):
    z = np.load(os.path.join(GOLDEN, 'race_%s.npz' % name))
    return {k: z[k] for k in z.files}


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def eq(a, b):
    """Bitwise-level equality for float tensors up to the sign of zero; NaN == NaN."""
    a = torch.as_tensor(a).cpu()
    b = torch.as_tensor(b).cpu()
    if a.shape != b.shape:
        return False
    if a.dtype.is_floating_point:
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    return bool((a == b).all())


def nmismatch(a, b):
    a = torch.as_tensor(a).cpu()
    b = torch.as_tensor(b).cpu()
    if a.dtype.is_floating_point:
        return int((~((a == b) | (torch.isnan(a) & torch.isnan(b)))).sum())
    return int((a != b).sum())
