"""Host mirror of ``isce3.core.make_projection(epsg).forward`` for the coordinate systems a
raster DEM can come in (cxx/isce3/core/Projections.cpp: createProj :373-402, UTM :84-213,
PolarStereo :247-297, CEA :324-360).  Used to build and sample synthetic DEMs on the host;
the device has its own implementation (csrc/projections.cuh) and the tests compare both with
the reference's compiled Projections.cpp."""
from __future__ import annotations

import math

A_WGS84 = 6378137.0
E2_WGS84 = 0.006694379990141317


def _clens(a, real):
    hr2, hr1 = 0.0, a[-1]
    c = 2.0 * math.cos(real)
    for ak in reversed(a[:-1]):
        hr2, hr1 = hr1, -hr2 + c * hr1 + ak
    return math.sin(real) * hr1


def _clenS(a, real, imag):
    sr, cr, sh, ch = math.sin(real), math.cos(real), math.sinh(imag), math.cosh(imag)
    r, im = 2.0 * cr * ch, -2.0 * sr * sh
    hr2 = hi2 = hi1 = 0.0
    hr1 = a[-1]
    for ak in reversed(a[:-1]):
        hr = -hr2 + r * hr1 - im * hi1 + ak
        hi = -hi2 + im * hr1 + r * hi1
        hr2, hi2, hr1, hi1 = hr1, hi1, hr, hi
    return sr * ch * hr1 - cr * sh * hi1, sr * ch * hi1 + cr * sh * hr1


def _tsfn(phi, sinphi, e):
    sinphi *= e
    return math.tan(0.5 * (0.5 * math.pi - phi)) / ((1.0 - sinphi) / (1.0 + sinphi)) ** (0.5 * e)


def _qsfn(sinphi, e, one_es):
    con = e * sinphi
    return one_es * (sinphi / (1.0 - con * con) - (0.5 / e) * math.log((1.0 - con) / (1.0 + con)))


class Projection:
    """forward(lon, lat) [radians] -> (x, y) in the CRS of ``epsg``."""

    def __init__(self, epsg: int):
        self.code = int(epsg)
        e2, a = E2_WGS84, A_WGS84
        if self.code == 4326:
            self.kind = "lonlat"
        elif 32600 < self.code <= 32660 or 32700 < self.code <= 32760:
            self.kind = "utm"
            self.isnorth = self.code <= 32660
            zone = self.code - (32600 if self.isnorth else 32700)
            self.lon0 = (zone - 0.5) * (math.pi / 30.0) - math.pi
            f = e2 / (1.0 + math.sqrt(1 - e2))
            n = f / (2.0 - f)
            self.cbg = [
                n * (-2 + n * ((2. / 3.) + n * ((4. / 3.) + n * ((-82. / 45.) + n * ((32. / 45.) + n * (4642. / 4725.)))))),
                n ** 2 * ((5. / 3.) + n * ((-16. / 15.) + n * ((-13. / 9.) + n * ((904. / 315.) + n * (-1522. / 945.))))),
                n ** 3 * ((-26. / 15.) + n * ((34. / 21.) + n * ((8. / 5.) + n * (-12686. / 2835.)))),
                n ** 4 * ((1237. / 630.) + n * ((-12. / 5.) + n * (-24832. / 14175.))),
                n ** 5 * ((-734. / 315.) + n * (109598. / 31185.)),
                n ** 6 * (444337. / 155925.)]
            self.Qn = (0.9996 / (1. + n)) * (1. + n * n * ((1. / 4.) + n * n * ((1. / 64.) + ((n * n) / 256.))))
            self.gtu = [
                n * (.5 + n * ((-2. / 3.) + n * ((5. / 16.) + n * ((41. / 180.) + n * ((-127. / 288.) + n * (7891. / 37800.)))))),
                n ** 2 * ((13. / 48.) + n * ((-3. / 5.) + n * ((557. / 1440.) + n * ((281. / 630.) + n * (-1983433. / 1935360.))))),
                n ** 3 * ((61. / 240.) + n * ((-103. / 140.) + n * ((15061. / 26880.) + n * (167603. / 181440.)))),
                n ** 4 * ((49561. / 161280.) + n * ((-179. / 168.) + n * (6601661. / 7257600.))),
                n ** 5 * ((34729. / 80640.) + n * (-3418889. / 1995840.)),
                n ** 6 * (212378941. / 319334400.)]
            Z = _clens(self.cbg, 0.0)
            self.Zb = -self.Qn * (Z + _clens(self.gtu, 2 * Z))
        elif self.code in (3031, 3413):
            self.kind = "polar"
            self.isnorth = self.code == 3413
            lat_ts = math.radians(70.0 if self.isnorth else 71.0)
            self.lon0 = math.radians(-45.0) if self.isnorth else 0.0
            self.e = math.sqrt(e2)
            self.akm1 = math.cos(lat_ts) / _tsfn(lat_ts, math.sin(lat_ts), self.e)
            self.akm1 *= a / math.sqrt(1.0 - self.e ** 2 * math.sin(lat_ts) ** 2)
        elif self.code == 6933:
            self.kind = "cea"
            lat_ts = math.pi / 6.0
            self.k0 = math.cos(lat_ts) / math.sqrt(1.0 - e2 * math.sin(lat_ts) ** 2)
            self.e = math.sqrt(e2)
            self.one_es = 1.0 - e2
        else:
            raise ValueError(f"Unknown EPSG code (in factory): {epsg}")

    def forward(self, lon: float, lat: float):
        a = A_WGS84
        if self.kind == "lonlat":
            return math.degrees(lon), math.degrees(lat)
        if self.kind == "utm":
            gauss = _clens(self.cbg, 2.0 * lat) + lat
            lam = lon - self.lon0
            Cn = math.atan2(math.sin(gauss), math.cos(lam) * math.cos(gauss))
            Ce = math.atan2(math.sin(lam) * math.cos(gauss),
                            math.hypot(math.sin(gauss), math.cos(gauss) * math.cos(lam)))
            Ce = math.asinh(math.tan(Ce))
            dCn, dCe = _clenS(self.gtu, 2 * Cn, 2 * Ce)
            Cn += dCn
            Ce += dCe
            if abs(Ce) > 2.623395162778:
                raise ValueError("point too far from the UTM central meridian")
            return (self.Qn * Ce * a + 500000.0,
                    (self.Qn * Cn + self.Zb) * a + (0.0 if self.isnorth else 10000000.0))
        if self.kind == "polar":
            sgn = 1.0 if self.isnorth else -1.0
            lam, phi = lon - self.lon0, lat * sgn
            temp = self.akm1 * _tsfn(phi, math.sin(phi), self.e)
            return temp * math.sin(lam), -temp * math.cos(lam) * sgn
        return (self.k0 * lon * a, 0.5 * a * _qsfn(math.sin(lat), self.e, self.one_es) / self.k0)


def make_projection(epsg: int) -> Projection:
    return Projection(epsg)


def utm_epsg_for(lon: float, lat: float) -> int:
    """EPSG code of the UTM zone holding (lon, lat) [radians]."""
    zone = int(math.floor((math.degrees(lon) + 180.0) / 6.0)) % 60 + 1
    return (32600 if lat >= 0 else 32700) + zone
