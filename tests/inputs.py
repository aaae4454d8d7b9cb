"""Encoders from oracle values to the two boundary layouts (test helper)."""
import numpy as np

from oracle import bigint_oracle as O


def points_le(points, nbytes):
    return b"".join(O.le_bytes(x, nbytes) + O.le_bytes(y, nbytes) for x, y in points)


def scalars_le(scalars):
    return b"".join(O.le_bytes(s, 32) for s in scalars)


def points_limb29(points, p, unreduced=None):
    mp = O.montgomery_params(p)
    return np.array(O.encode_affine_limb29(points, mp, unreduced), dtype=np.uint32)


def te_points_limb29(points, p):
    mp = O.montgomery_params(p)
    return np.array(O.encode_te_limb29(points, mp), dtype=np.uint32)


def scalars_limb29(scalars, q):
    mp = O.montgomery_params(q, 29, 1)
    return np.array(O.encode_scalars_limb29(scalars, mp), dtype=np.uint32)
