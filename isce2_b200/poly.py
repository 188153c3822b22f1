"""Host-side Poly2D / Poly1D with the interface of isceobj.Util.Poly2D / Poly1D that the zero-Doppler
components use (components/isceobj/Util/Library/python/Poly2D.py:131-265, Poly1D.py:100-215).  A real ISCE
polynomial object is accepted wherever one of these is (duck-typed getters)."""
from __future__ import annotations


class Poly2D:
    """coeffs[azimuth][range]; value = sum_ij c[i][j] * ((azi-meanAz)/normAz)^i * ((rng-meanRg)/normRg)^j"""

    def __init__(self, family="", name=""):
        self.family = family
        self.name = name
        self._width = None
        self._length = None
        self._rangeOrder = None
        self._azimuthOrder = None
        self._normRange = 1.0
        self._meanRange = 0.0
        self._normAzimuth = 1.0
        self._meanAzimuth = 0.0
        self._coeffs = []

    def initPoly(self, rangeOrder=None, azimuthOrder=None, coeffs=None, image=None):
        if coeffs:
            import copy
            self._coeffs = copy.deepcopy([list(map(float, row)) for row in coeffs])
        self._rangeOrder = int(rangeOrder)
        self._azimuthOrder = int(azimuthOrder)
        if image is not None:
            self._width = image.width
            self._length = image.length

    def setCoeffs(self, parms):
        self._coeffs = [list(map(float, row)) for row in parms]

    def getCoeffs(self): return self._coeffs
    def setWidth(self, v): self._width = int(v)
    def setLength(self, v): self._length = int(v)
    def getWidth(self): return self._width
    def getLength(self): return self._length
    def setNormRange(self, v): self._normRange = float(v)
    def setMeanRange(self, v): self._meanRange = float(v)
    def getNormRange(self): return self._normRange
    def getMeanRange(self): return self._meanRange
    def setNormAzimuth(self, v): self._normAzimuth = float(v)
    def setMeanAzimuth(self, v): self._meanAzimuth = float(v)
    def getNormAzimuth(self): return self._normAzimuth
    def getMeanAzimuth(self): return self._meanAzimuth
    def getRangeOrder(self): return self._rangeOrder
    def getAzimuthOrder(self): return self._azimuthOrder
    width = property(getWidth, setWidth)
    length = property(getLength, setLength)

    def __call__(self, azi, rng):
        y = (azi - self._meanAzimuth) / self._normAzimuth
        x = (rng - self._meanRange) / self._normRange
        res, sy = 0.0, 1.0
        for row in self._coeffs:
            sx = 1.0
            for c in row:
                res += sx * sy * c
                sx *= x
            sy *= y
        return res

    def copy(self):
        import copy
        return copy.deepcopy(self)

    # no C pointers behind this implementation; kept so that code written for the reference keeps working
    def createPoly2D(self): return None
    def getPointer(self): return 0
    def finalize(self): return None
    def finalizeImage(self): return None


class Poly1D:
    def __init__(self, family="", name="", order=None, image=None, direction="x"):
        self.family = family
        self.name = name
        self._order = order
        self._norm = 1.0
        self._mean = 0.0
        self._coeffs = []
        self._width = None
        self._length = None

    def initPoly(self, order=None, coeffs=None, image=None, direction="x"):
        if coeffs is not None:
            self._coeffs = [float(c) for c in coeffs]
        self._order = int(order)

    def setCoeffs(self, parms): self._coeffs = [float(c) for c in parms]
    def getCoeffs(self): return self._coeffs
    def setNorm(self, v): self._norm = float(v)
    def setMean(self, v): self._mean = float(v)
    def getNorm(self): return self._norm
    def getMean(self): return self._mean
    def getOrder(self): return self._order

    def __call__(self, rng):
        x = (rng - self._mean) / self._norm
        res, sx = 0.0, 1.0
        for c in self._coeffs:
            res += sx * c
            sx *= x
        return res

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def exportToC(self): return 0
    def createPoly1D(self): return None
    def finalize(self): return None


def poly2d_fields(p):
    """(coeffs 2-D list, meanRange, meanAzimuth, normRange, normAzimuth) of ours or of an ISCE Poly2D."""
    coeffs = p.getCoeffs() if hasattr(p, "getCoeffs") else p._coeffs
    return ([list(map(float, r)) for r in coeffs], float(p.getMeanRange()), float(p.getMeanAzimuth()),
            float(p.getNormRange()), float(p.getNormAzimuth()))


def poly1d_fields(p):
    coeffs = p.getCoeffs() if hasattr(p, "getCoeffs") else p._coeffs
    return [float(c) for c in coeffs], float(p.getMean()), float(p.getNorm())
