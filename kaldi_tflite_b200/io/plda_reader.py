"""
Kaldi PLDA model reader (binary): `<Plda>` mean (vector), transform (matrix),
psi (vector) `</Plda>`.  Same surface as the reference's
`io/kaldi/plda_reader.py:22-62` (`.mean`, `.transformMat`, `.psi`).
"""

from .object_reader import KaldiObjReader


class KaldiPldaReader(KaldiObjReader):

    def __init__(self, plda_path: str, binary: bool):
        super().__init__(plda_path, binary)
        self.mean = None
        self.transformMat = None
        self.psi = None
        self.read()

    def read(self):
        self.expectToken("<Plda>")
        self.mean = self.readVec()
        self.transformMat = self.readMat()
        self.psi = self.readVec()
        self.expectToken("</Plda>")
