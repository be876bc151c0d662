"""
Reader for the primitive encodings of Kaldi binary objects.

Host-side, load-time only (weights are parsed once and packed into device
buffers).  Mirrors the public surface of the reference's
`kaldi_tflite/lib/io/kaldi/object_reader.py:23-483` (`KaldiObjReader` and its
`read*/expect*/peek*` methods) so that code written against the reference
keeps working; the implementation is an independent cursor over a
`memoryview`.

Encodings (Kaldi `base/io-funcs`, `matrix/kaldi-vector.cc`, `kaldi-matrix.cc`):
  basic types   <size:1 byte><little-endian value>
  bool          'T' | 'F'
  vector        "FV " | "DV "  + \\x04 + int32 dim + data
  matrix        "FM " | "DM "  + \\x04 + int32 rows + \\x04 + int32 cols + row-major data
  packed sym.   "FP " | "DP "  + \\x04 + int32 rows + lower-triangular data
"""

from typing import Iterable, Union

import numpy as np

_VEC_TYPES = {"FV ": np.float32, "DV ": np.float64}
_MAT_TYPES = {"FM ": np.float32, "DM ": np.float64}
_PACKED_TYPES = {"FP ": np.float32, "DP ": np.float64}


class KaldiObjReader:

    def __init__(self, path: str, binary: bool):
        if not binary:
            raise NotImplementedError("objects in text format are currently not supported")
        self.path = path
        self.binary = binary
        self.curPos = 0
        with open(path, "rb") as fd:
            self.data = fd.read()

    # -- raw byte access -------------------------------------------------
    def readBytes(self, nBytes: int) -> bytes:
        if self.curPos >= len(self.data):
            return b""
        buf = self.data[self.curPos:self.curPos + nBytes]
        self.curPos += len(buf)
        return buf

    def peekBytes(self, nBytes: int) -> bytes:
        if self.curPos >= len(self.data):
            return b""
        return self.data[self.curPos:self.curPos + nBytes]

    # -- text-ish helpers ------------------------------------------------
    def expectLine(self):
        nl = self.data.find(b"\n", self.curPos)
        if nl < 0:
            raise ValueError("expected new line but did not get any")
        self.curPos = nl + 1

    def readLine(self) -> str:
        nl = self.data.find(b"\n", self.curPos)
        if nl < 0:
            raise ValueError("expected new line but did not get any")
        line = self.data[self.curPos:nl].decode()
        self.curPos = nl + 1
        return line

    def expectToken(self, token: str, stopTokens: Iterable[str] = ()) -> bool:
        """
        Scans forward for `token`.  On success the cursor moves just past the
        token and its trailing separator and True is returned.  If one of
        `stopTokens` occurs earlier, returns False without moving.  Raises
        ValueError when neither is found.
        """
        tok = token.encode("utf-8")
        hit = self.data.find(tok, self.curPos)
        firstStop = -1
        for s in stopTokens:
            p = self.data.find(s.encode("utf-8"), self.curPos)
            if p >= 0 and (firstStop < 0 or p < firstStop):
                firstStop = p
        if hit >= 0 and (firstStop < 0 or hit <= firstStop):
            self.curPos = hit + len(tok) + 1
            return True
        if firstStop >= 0:
            return False
        raise ValueError(f"failed to find expected token '{token}'")

    def readToken(self) -> str:
        pos = self.curPos
        while True:
            sp = self.data.find(b" ", pos)
            if sp < 0:
                raise ValueError(f"no whitespace separated token after pos {self.curPos}")
            try:
                token = self.data[self.curPos:sp].decode()
            except UnicodeDecodeError:
                pos = sp + 1
                continue
            self.curPos = sp + 1
            return token

    # -- basic types -------------------------------------------------------
    def readBasicType(self, dtype) -> Union[int, float]:
        want = np.dtype(dtype).itemsize
        got = int.from_bytes(self.readBytes(1), "little")
        if got != want:
            raise ValueError(
                f"data type read is specified using {got} bytes, but want to parse {want} bytes")
        buf = self.readBytes(got)
        if len(buf) != want:
            raise ValueError(f"failed to parse any value of type {dtype}")
        return np.frombuffer(buf, dtype=dtype)[0]

    def readInt(self) -> int:
        return self.readBasicType(np.int32)

    def readFloat(self) -> float:
        return self.readBasicType(np.float32)

    def readDouble(self) -> float:
        return self.readBasicType(np.float64)

    def readBool(self) -> bool:
        b = self.readBytes(1)
        if b == b"T":
            return True
        if b == b"F":
            return False
        raise ValueError(f"unexpected format for booleans, expected 'T' or 'F', got {b}")

    # -- arrays ------------------------------------------------------------
    def _header(self, table, what):
        header = self.readBytes(3).decode(errors="replace")
        if header not in table:
            raise ValueError(f"unknown header for {what} type '{header}'")
        return table[header]

    def readVec(self) -> np.ndarray:
        dt = self._header(_VEC_TYPES, "vector")
        dim = int(self.readInt())
        if dim == 0:
            return np.array([], dtype=dt)
        buf = self.readBytes(dim * np.dtype(dt).itemsize)
        return np.frombuffer(buf, dtype=dt)

    def readMat(self) -> np.ndarray:
        if self.peekBytes(2) == b"CM":
            raise NotImplementedError("can't decode compressed matrix yet")
        dt = self._header(_MAT_TYPES, "matrix")
        rows = int(self.readInt())
        cols = int(self.readInt())
        if rows == 0 or cols == 0:
            return np.zeros((rows, cols), dtype=dt)
        buf = self.readBytes(rows * cols * np.dtype(dt).itemsize)
        return np.frombuffer(buf, dtype=dt).reshape(rows, cols)

    def readPackedMat(self) -> np.ndarray:
        dt = self._header(_PACKED_TYPES, "matrix")
        rows = int(self.readInt())
        if rows == 0:
            return np.zeros((0, 0), dtype=dt)
        n = ((rows + 1) * rows) // 2
        tri = np.frombuffer(self.readBytes(n * np.dtype(dt).itemsize), dtype=dt)
        full = np.zeros((rows, rows), dtype=dt)
        il = np.tril_indices(rows)
        full[il] = tri
        full.T[il] = tri
        return full
