"""merge_arrays, the helper behind the cumulative load (reference utils.py:26-34,
called at streets4mpi.py:100).  Host-side convenience for ``array('I')`` inputs;
the device-resident equivalent is ``s4_sim_commit_total``."""
from array import array

import numpy as np


def merge_arrays(arrays):
    """Element-wise sum of ``array('I')`` operands, ``None`` entries skipped.

    Like the reference, the result is a fresh ``array('I')`` sized after the first
    operand, and a sum that does not fit 32 bits raises ``OverflowError`` (that is
    what ``array('I').__setitem__`` did there)."""
    n = len(arrays[0])
    total = np.zeros(n, dtype=np.uint64)
    for arr in arrays:
        if arr is None:
            continue
        part = np.frombuffer(arr, dtype=np.uint32) if isinstance(arr, array) else np.asarray(arr, dtype=np.uint32)
        total[:len(part)] += part
    if n and int(total.max()) > 0xFFFFFFFF:
        raise OverflowError("unsigned int is greater than maximum")
    out = array("I")
    out.frombytes(total.astype(np.uint32).tobytes())
    return out
