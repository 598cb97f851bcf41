"""tf.linalg subset for the shim (TEST INFRASTRUCTURE, see tensorflow/__init__.py)."""
import torch as _t


def band_part(x, num_lower, num_upper):
    """tf.linalg.band_part: keep entries with (num_lower < 0 or i-j <= num_lower) and
    (num_upper < 0 or j-i <= num_upper) of the innermost two dims."""
    m, n = x.shape[-2], x.shape[-1]
    i = _t.arange(m).unsqueeze(1)
    j = _t.arange(n).unsqueeze(0)
    keep = _t.ones(m, n, dtype=_t.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return x * keep.to(x.dtype)
