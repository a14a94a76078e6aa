// Map projections of raster DEMs: forward (lon, lat) -> (x, y) of the DEM's CRS, as
// DEMInterpolator::interpolateLonLat needs it (cxx/isce3/geometry/DEMInterpolator.cpp:592-611).
// Behavioural reference: cxx/isce3/core/Projections.cpp (LonLat :Projections.h:127-133,
// UTM :84-240 Krueger series as in PROJ's etmerc, PolarStereo :247-318, CEA :324-371) and
// the factory createProj :373-402.  Parameters are set up on the host once per call
// (proj_setup) and travel inside the DEM descriptor; the device only runs proj_forward.
#pragma once
#include <cmath>

namespace i3b {

enum { PROJ_LONLAT = 0, PROJ_UTM = 1, PROJ_POLAR = 2, PROJ_CEA = 3 };

struct DevProj {
    int kind, isnorth;
    double a;            // semi-major axis
    double lon0;         // central meridian (UTM, polar stereographic)
    double Qn, Zb;       // UTM: normalised meridian quadrant, origin offset
    double cbg[6], gtu[6];
    double akm1, e;      // polar stereographic scale, eccentricity
    double k0, one_es;   // cylindrical equal area
};

namespace projdetail {

__host__ __device__ inline double clens(const double* a, int size, double real)
{
    // Clenshaw summation of sum a_k sin(2 k B) (Projections.cpp:36-46)
    double hr = 0., hr1 = a[size - 1], hr2 = 0.;
    const double c = 2. * cos(real);
    for (int i = size - 2; i >= 0; --i) {
        hr = -hr2 + c * hr1 + a[i];
        hr2 = hr1;
        hr1 = hr;
    }
    return sin(real) * hr1;
}

// complex Clenshaw summation (Projections.cpp:58-82)
__host__ __device__ inline void clenS(const double* a, int size, double real, double imag, double* R,
                                      double* I)
{
    const double sr = sin(real), cr = cos(real), sh = sinh(imag), ch = cosh(imag);
    const double r = 2. * cr * ch, im = -2. * sr * sh;
    double hr = 0., hr1 = a[size - 1], hr2 = 0., hi = 0., hi1 = 0., hi2 = 0.;
    for (int k = size - 2; k >= 0; --k) {
        hr = -hr2 + r * hr1 - im * hi1 + a[k];
        hi = -hi2 + im * hr1 + r * hi1;
        hr2 = hr1; hi2 = hi1;
        hr1 = hr; hi1 = hi;
    }
    *R = sr * ch * hr1 - cr * sh * hi1;
    *I = sr * ch * hi1 + cr * sh * hr1;
}

__host__ __device__ inline double pj_tsfn(double phi, double sinphi, double e)
{
    sinphi *= e;
    return tan(.5 * (.5 * M_PI - phi)) / pow((1. - sinphi) / (1. + sinphi), .5 * e);
}

__host__ __device__ inline double pj_qsfn(double sinphi, double e, double one_es)
{
    const double con = e * sinphi;
    return one_es * ((sinphi / (1. - con * con)) - ((.5 / e) * log((1. - con) / (1. + con))));
}

} // namespace projdetail

// Host: parameters for an EPSG code; false when createProj would not know it as a DEM CRS
// (4978 geocentric is a valid isce3 projection but not a raster coordinate system).
inline bool proj_setup(int epsg, double a, double e2, DevProj* p)
{
    *p = DevProj {};
    p->a = a;
    if (epsg == 4326) {
        p->kind = PROJ_LONLAT;
        return true;
    }
    if (epsg > 32600 && epsg < 32800) {
        int zone;
        if (epsg <= 32660) {
            zone = epsg - 32600;
            p->isnorth = 1;
        } else if (epsg > 32700 && epsg <= 32760) {
            zone = epsg - 32700;
            p->isnorth = 0;
        } else {
            return false;
        }
        p->kind = PROJ_UTM;
        p->lon0 = ((zone - 0.5) * (M_PI / 30.)) - M_PI;
        const double f = e2 / (1. + std::sqrt(1 - e2));
        const double n = f / (2. - f);
        double* cbg = p->cbg;
        double* gtu = p->gtu;
        cbg[0] = n * (-2 + n * ((2. / 3.) + n * ((4. / 3.) + n * ((-82. / 45.) + n * ((32. / 45.) + n * (4642. / 4725.))))));
        cbg[1] = std::pow(n, 2) * ((5. / 3.) + n * ((-16. / 15.) + n * ((-13. / 9.) + n * ((904. / 315.) + n * (-1522. / 945.)))));
        cbg[2] = std::pow(n, 3) * ((-26. / 15.) + n * ((34. / 21.) + n * ((8. / 5.) + n * (-12686. / 2835.))));
        cbg[3] = std::pow(n, 4) * ((1237. / 630.) + n * ((-12. / 5.) + n * (-24832. / 14175.)));
        cbg[4] = std::pow(n, 5) * ((-734. / 315.) + n * (109598. / 31185.));
        cbg[5] = std::pow(n, 6) * (444337. / 155925.);
        p->Qn = (0.9996 / (1. + n)) * (1. + n * n * ((1. / 4.) + n * n * ((1. / 64.) + ((n * n) / 256.))));
        gtu[0] = n * (.5 + n * ((-2. / 3.) + n * ((5. / 16.) + n * ((41. / 180.) + n * ((-127. / 288.) + n * (7891. / 37800.))))));
        gtu[1] = std::pow(n, 2) * ((13. / 48.) + n * ((-3. / 5.) + n * ((557. / 1440.) + n * ((281. / 630.) + n * (-1983433. / 1935360.)))));
        gtu[2] = std::pow(n, 3) * ((61. / 240.) + n * ((-103. / 140.) + n * ((15061. / 26880.) + n * (167603. / 181440.))));
        gtu[3] = std::pow(n, 4) * ((49561. / 161280.) + n * ((-179. / 168.) + n * (6601661. / 7257600.)));
        gtu[4] = std::pow(n, 5) * ((34729. / 80640.) + n * (-3418889. / 1995840.));
        gtu[5] = std::pow(n, 6) * (212378941. / 319334400.);
        const double Z = projdetail::clens(cbg, 6, 0.);
        p->Zb = -p->Qn * (Z + projdetail::clens(gtu, 6, 2 * Z));
        return true;
    }
    if (epsg == 3031 || epsg == 3413) {
        p->kind = PROJ_POLAR;
        double lat_ts;
        if (epsg == 3031) {
            p->isnorth = 0;
            lat_ts = (71. * M_PI) / 180.;
            p->lon0 = 0.;
        } else {
            p->isnorth = 1;
            lat_ts = 70. * (M_PI / 180.);
            p->lon0 = -45. * (M_PI / 180.);
        }
        p->e = std::sqrt(e2);
        p->akm1 = std::cos(lat_ts) / projdetail::pj_tsfn(lat_ts, std::sin(lat_ts), p->e);
        p->akm1 *= a / std::sqrt(1. - (std::pow(p->e, 2) * std::pow(std::sin(lat_ts), 2)));
        return true;
    }
    if (epsg == 6933) {
        p->kind = PROJ_CEA;
        const double lat_ts = M_PI / 6.;
        p->k0 = std::cos(lat_ts) / std::sqrt(1. - (e2 * std::pow(std::sin(lat_ts), 2)));
        p->e = std::sqrt(e2);
        p->one_es = 1. - e2;
        return true;
    }
    return false;
}

// (lon, lat) [rad] -> (x, y) of the CRS.  Returns nonzero where the reference's forward()
// does (UTM too far from the central meridian); PROJ_LONLAT gives degrees.
__host__ __device__ inline int proj_forward(const DevProj& p, double lon, double lat, double* x, double* y)
{
    switch (p.kind) {
    case PROJ_UTM: {
        const double gauss = projdetail::clens(p.cbg, 6, 2. * lat) + lat;
        const double lam = lon - p.lon0;
        const double sg = sin(gauss), cg = cos(gauss), sl = sin(lam), cl = cos(lam);
        double Cn = atan2(sg, cl * cg);
        double Ce = atan2(sl * cg, hypot(sg, cg * cl));
        Ce = asinh(tan(Ce));
        double dCn, dCe;
        projdetail::clenS(p.gtu, 6, 2 * Cn, 2 * Ce, &dCn, &dCe);
        Cn += dCn;
        Ce += dCe;
        if (fabs(Ce) > 2.623395162778) return 1;
        *x = (p.Qn * Ce * p.a) + 500000.;
        *y = (((p.Qn * Cn) + p.Zb) * p.a) + (p.isnorth ? 0. : 10000000.);
        return 0;
    }
    case PROJ_POLAR: {
        const double sgn = p.isnorth ? 1. : -1.;
        const double lam = lon - p.lon0;
        const double phi = lat * sgn;
        const double temp = p.akm1 * projdetail::pj_tsfn(phi, sin(phi), p.e);
        *x = temp * sin(lam);
        *y = -temp * cos(lam) * sgn;
        return 0;
    }
    case PROJ_CEA:
        *x = p.k0 * lon * p.a;
        *y = (.5 * p.a * projdetail::pj_qsfn(sin(lat), p.e, p.one_es)) / p.k0;
        return 0;
    default:
        *x = lon * 180.0 / M_PI;
        *y = lat * 180.0 / M_PI;
        return 0;
    }
}

} // namespace i3b
