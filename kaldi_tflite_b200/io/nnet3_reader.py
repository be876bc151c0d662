"""
Kaldi nnet3 `final.raw` reader (binary).  Load-time only.

Public surface follows the reference's `io/kaldi/nnet3_reader.py:27-329`:
`KaldiNnet3Reader(path, binary)` with `.config` (list of config lines),
`.components` (list of dicts with "name", "type" and the parsed fields),
`getComponent(pattern)` and `getWeights(pattern)`; names are matched with
`re.match(pattern, component_name)` exactly like the reference (:282), so
`"tdnn1.affine"` also selects by prefix.
"""

import re
from typing import Iterable, List

import numpy as np

from .object_reader import KaldiObjReader

_NONLINEAR = {"Sigmoid", "Tanh", "RectifiedLinear", "Softmax", "LogSoftmax", "NoOp"}


class KaldiNnet3Reader(KaldiObjReader):

    def __init__(self, nnet3_path: str, binary: bool):
        super().__init__(nnet3_path, binary)
        self.config: List[str] = []
        self.components: List[dict] = []
        self.read()

    # field tables: (token, reader-method-name, key)
    def getComponentFormat(self, compType: str) -> list:
        comp = self.stripTagsAndSuffix(compType, suffix="Component")
        if comp in _NONLINEAR:
            return [("<Dim>", self.readInt, "dim"),
                    ("<ValueAvg>", self.readVec, "value-avg"),
                    ("<DerivAvg>", self.readVec, "deriv-avg"),
                    ("<Count>", self.readDouble, "count"),
                    ("<OderivRms>", self.readVec, "oderiv-rms"),
                    ("<OderivCount>", self.readDouble, "oderiv-count")]
        if comp in {"Affine", "NaturalGradientAffine"}:
            return [("<LinearParams>", self.readMat, "params"),
                    ("<BiasParams>", self.readVec, "bias")]
        if comp == "Linear":
            return [("<Params>", self.readMat, "params")]
        if comp == "BatchNorm":
            return [("<Dim>", self.readInt, "dim"),
                    ("<BlockDim>", self.readInt, "block-dim"),
                    ("<Epsilon>", self.readFloat, "epsilon"),
                    ("<TargetRms>", self.readFloat, "target-rms"),
                    ("<TestMode>", self.readBool, "test-mode"),
                    ("<Count>", self.readDouble, "count"),
                    ("<StatsMean>", self.readVec, "stats-mean"),
                    ("<StatsVar>", self.readVec, "stats-var")]
        if comp in {"StatisticsExtraction", "StatisticsPooling"}:
            return []
        raise ValueError(f"unsupported component type '{compType}'")

    @staticmethod
    def stripTagsAndSuffix(token: str, suffix: str = "") -> str:
        token = token.lstrip("<")
        if token.endswith("/>"):
            token = token.rstrip("/>")
        if token.endswith(">"):
            token = token.rstrip(">")
        if suffix and token.endswith(suffix):
            token = token[:-len(suffix)]
        return token

    def read(self):
        self.expectToken("<Nnet3>")
        if self.readLine().strip() != "":
            raise ValueError("expected model config following <Nnet3> token, got blank line")
        self.readConfigLines()

        self.expectToken("<NumComponents>")
        n = int(self.readInt())
        assert 0 < n < 100000, f"expected between 1 and 9999 components, got {n}"

        self.components = []
        for _ in range(n):
            self.expectToken("<ComponentName>")
            name = self.readToken()
            ctype = self.readToken()
            comp = {"name": name, "type": ctype}
            comp.update(self.readComponent(ctype))
            self.components.append(comp)
        self.expectToken("</Nnet3>")

    def readConfigLines(self):
        self.config = []
        line = self.readLine().strip()
        while line != "":
            self.config.append(line)
            line = self.readLine().strip()

    def readComponent(self, compType: str) -> dict:
        closing = {"</" + compType[1:], "<ComponentName>"}
        out = {}
        for token, fn, key in self.getComponentFormat(compType):
            if self.expectToken(token, closing):
                out[key] = fn()
            else:
                print(f"  - failed to find token {token}")
        return out

    def getComponent(self, name: str) -> Iterable[dict]:
        return [c for c in self.components
                if c.get("name") is not None and re.match(rf"{name}", c["name"])]

    def getWeights(self, name: str) -> Iterable[np.ndarray]:
        matching = self.getComponent(name)
        if len(matching) == 0:
            raise KeyError(f"no components with name matching '{name}'")
        weights = []
        for c in matching:
            if c["type"] == "<NaturalGradientAffineComponent>":
                weights.extend([c["params"], c["bias"]])
            elif c["type"] == "<BatchNormComponent>":
                weights.extend([c["target-rms"], c["stats-mean"], c["stats-var"]])
        return weights
