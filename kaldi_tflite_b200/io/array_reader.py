"""
Reader for stand-alone Kaldi `.mat` / `.vec` files, binary (`\\0B` + FM/DM/FV/DV
object) or text (`[ ... ]`).  Same call signature and error behaviour as the
reference's `io/kaldi/array_reader.py:24-104`.
"""

import numpy as np

from .object_reader import KaldiObjReader

_FLOATS = (np.float32, np.float64)
_INTS = (np.int16, np.int32, np.int64)


def ReadKaldiArray(path: str, binary: bool, dtype=np.float32) -> np.ndarray:
    if binary:
        r = KaldiObjReader(path, True)
        r.readBytes(2)                       # "\0B" binary marker
        kind = r.peekBytes(2).decode(errors="replace")
        if kind in ("FM", "DM", "CM"):
            return r.readMat()
        if kind in ("FV", "DV"):
            return r.readVec()
        raise ValueError(
            f"binary file contains unexpected header bytes, {kind}, "
            "expected 'FV', 'DV', 'FM', 'DM' or 'CM'")

    if dtype in _FLOATS:
        conv = float
    elif dtype in _INTS:
        conv = int
    else:
        raise ValueError(f"unsupported data type: {dtype}")

    rows = []
    with open(path, "r") as f:
        for line in f:
            toks = line.split()
            opened = "[" in toks
            closed = "]" in toks
            vals = [conv(t) for t in toks if t not in ("[", "]")]
            if opened and closed:            # vector on a single line
                return np.array(vals, dtype=dtype)
            if vals:
                rows.append(vals)
            if closed:
                return np.array(rows, dtype=dtype)
    raise ValueError("reached end of file without finding closing bracket for matrix")
