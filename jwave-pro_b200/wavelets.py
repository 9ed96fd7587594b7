"""Wavelet objects: coefficient carriers with the getters of transforms/wavelets/Wavelet.java:152-219.

Tables come from filters_generated.json (tools/extract_filters.py, data only); the other three filters are built by
the rule of Wavelet._buildOrthonormalSpace (Wavelet.java:104-122).  Scope: Haar1, Daubechies2..20, Symlet2..20,
Coiflet1..5 -- the families BASELINE.json:north_star names.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "filters_generated.json")) as _f:
    _TABLES = {t["class"]: t for t in json.load(_f)["wavelets"]}


class Wavelet:
    """Orthonormal wavelet: scalingDeCom given, waveletDeCom[i] = (-1)^i scalingDeCom[L-1-i], recon = decom."""

    def __init__(self, cls_name=None, scaling=None, name=None):
        if cls_name is not None:
            t = _TABLES[cls_name]
            scaling, name = t["scalingDeCom"], t["name"]
        self._name = name
        self._scalingDeCom = np.array(scaling, dtype=np.float64)
        L = len(self._scalingDeCom)
        self._motherWavelength = L
        self._transformWavelength = 2
        w = np.empty(L)
        for i in range(L):
            w[i] = self._scalingDeCom[L - 1 - i] if i % 2 == 0 else -self._scalingDeCom[L - 1 - i]
        self._waveletDeCom = w
        self._scalingReCon = self._scalingDeCom.copy()
        self._waveletReCon = self._waveletDeCom.copy()

    # getters return copies, like the reference (Wavelet.java:178-219)
    def getName(self):
        return self._name

    def getMotherWavelength(self):
        return self._motherWavelength

    def getTransformWavelength(self):
        return self._transformWavelength

    def getScalingDeComposition(self):
        return self._scalingDeCom.copy()

    def getWaveletDeComposition(self):
        return self._waveletDeCom.copy()

    def getScalingReConstruction(self):
        return self._scalingReCon.copy()

    def getWaveletReConstruction(self):
        return self._waveletReCon.copy()

    def __repr__(self):
        return "Wavelet(%r, L=%d)" % (self._name, self._motherWavelength)


def _make(cls_name):
    def ctor():
        return Wavelet(cls_name)
    ctor.__name__ = cls_name
    ctor.__doc__ = "wavelets/%s/%s.java" % (_TABLES[cls_name]["package"], cls_name)
    return ctor


ALL_CLASSES = list(_TABLES.keys())
for _c in ALL_CLASSES:
    globals()[_c] = _make(_c)


def create(name):
    """Factory by class name ("Daubechies4") or display name ("Daubechies 4"), cf. WaveletBuilder.create."""
    if name in _TABLES:
        return Wavelet(name)
    for c, t in _TABLES.items():
        if t["name"] == name:
            return Wavelet(c)
    raise KeyError(name)


def create2arr():
    """The orthogonal part of WaveletBuilder.create2arr() (WaveletBuilder.java:427-502) that is in scope."""
    return [Wavelet(c) for c in ALL_CLASSES]
