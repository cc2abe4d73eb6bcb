// bsg_math.cuh -- device-side float32 restatement of the BlueSky scalar maths on the step path.
// Mirrors (does not copy) bluesky/tools/aero.py and geo.py as restated in oracle/aero.py, oracle/geo.py;
// reference call sites: every bs.traf.cre / bs.sim.step() (e.g. horizontal_cr_env.py:91,109) and the
// bs.tools.geo.kwik* calls of the env files (e.g. horizontal_cr_env.py:169,190,265).
// Positions are float64 (lat/lon accumulate in double); everything else is float32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsg {

constexpr float kKts = 0.514444f;
constexpr float kFt = 0.3048f;
constexpr float kFpm = 0.3048f / 60.0f;
constexpr float kNm = 1852.0f;
constexpr float kG0 = 9.80665f;
constexpr float kRgas = 287.05287f;
constexpr float kP0 = 101325.0f;
constexpr float kRho0 = 1.225f;
constexpr float kT0 = 288.15f;
constexpr float kTstrat = 216.65f;
constexpr float kGammaR = 1.4f * 287.05287f;
constexpr float kRearth = 6371000.0f;
constexpr double kRearthD = 6371000.0;
constexpr float kDeg2Rad = 0.017453292519943295f;
constexpr float kRad2Deg = 57.29577951308232f;
constexpr double kDeg2RadD = 0.017453292519943295;
constexpr double kRad2DegD = 57.29577951308232;
constexpr float kTanBankDef = 0.4663076581549986f;   // tan(25 deg), Autopilot.bankdef
constexpr float kVsDef = 1500.0f * kFpm;             // Autopilot.vsdef
constexpr float kAzMax = 300.0f * kFpm;

// x^y for x > 0 through the accurate log2f / exp2f (no fast-math): ~1e-7 relative.
__device__ __forceinline__ float powpos(float x, float y) { return exp2f(y * log2f(x)); }

struct Atmos { float p, rho, T; };

__device__ __forceinline__ Atmos vatmos(float h) {
    Atmos a;
    a.T = fmaxf(kT0 - 0.0065f * h, kTstrat);
    float rhotrop = kRho0 * powpos(a.T * (1.0f / kT0), 4.256848030018761f);
    float dh = fmaxf(0.0f, h - 11000.0f);
    a.rho = (dh > 0.0f) ? rhotrop * expf(-dh * (1.0f / 6341.552161f)) : rhotrop;
    a.p = a.rho * kRgas * a.T;
    return a;
}

__device__ __forceinline__ float vsound(const Atmos& a) { return sqrtf(kGammaR * a.T); }

// (1 + x)^y - 1 without the cancellation of pow(1 + x, y) - 1: the CAS<->TAS round trips feed back into
// each other every env step (selspd <- cas + dv <- tas2cas(tas)), so their error must stay ~1 ulp.
__device__ __forceinline__ float pow1pm1(float x, float y) { return expm1f(y * log1pf(x)); }

__device__ __forceinline__ float tas2cas(float tas, const Atmos& a) {
    float q = a.p * pow1pm1(a.rho * tas * tas / (7.0f * a.p), 3.5f);
    float c = sqrtf(7.0f * kP0 / kRho0 * pow1pm1(q * (1.0f / kP0), 2.0f / 7.0f));
    return tas < 0.0f ? -c : c;
}

__device__ __forceinline__ float cas2tas(float cas, const Atmos& a) {
    float q = kP0 * pow1pm1(kRho0 * cas * cas / (7.0f * kP0), 3.5f);
    float t = sqrtf(7.0f * a.p / a.rho * pow1pm1(q / a.p, 2.0f / 7.0f));
    return cas < 0.0f ? -t : t;
}

// vcasormach2tas: |spd| < 1 is a Mach number
__device__ __forceinline__ float casormach2tas(float spd, const Atmos& a) {
    return fabsf(spd) < 1.0f ? spd * vsound(a) : cas2tas(spd, a);
}

// numpy's a % 360 (result in [0, 360)); floor form instead of fmodf (~6 instructions instead of ~40)
__device__ __forceinline__ float mod360(float a) {
    float r = fmaf(-360.0f, floorf(a * (1.0f / 360.0f)), a);
    r = r < 0.0f ? r + 360.0f : r;
    return r >= 360.0f ? r - 360.0f : r;
}
// (a + 180) % 360 - 180
__device__ __forceinline__ float degto180(float a) { return mod360(a + 180.0f) - 180.0f; }
// functions.py:4-22: a single +-360 fold with strict inequalities
__device__ __forceinline__ float wrap180_fold(float a) {
    return a > 180.0f ? a - 360.0f : (a < -180.0f ? a + 360.0f : a);
}
__device__ __forceinline__ double dmod360(double a) {
    double r = fmod(a, 360.0);
    return r < 0.0 ? r + 360.0 : r;
}

// geo.kwikqdrdist: differences in float64, the rest in float32.  qdr [deg 0..360), dist [NM].
__device__ __forceinline__ void kwikqdrdist(double lata, double lona, double latb, double lonb,
                                            float& qdr, float& dist_nm) {
    float dlat = (float)((latb - lata) * kDeg2RadD);
    float dlon = (float)((dmod360((lonb - lona) + 180.0) - 180.0) * kDeg2RadD);
    float cavelat = __cosf((float)((lata + latb) * (0.5 * kDeg2RadD)));      // |mean lat| <= pi/2: ~4e-7 absolute
    float dx = dlon * cavelat;
    dist_nm = (kRearth / kNm) * sqrtf(dlat * dlat + dx * dx);
    qdr = mod360(kRad2Deg * atan2f(dx, dlat));
}

// geo.kwikqdrdist without the bearing angle: the flat-earth offsets (radians of arc, north / east) and the distance [NM].
// cos / sin of the bearing are dn / ang, de / ang: what an observation needs when it only takes the cosine and sine of
// (heading - bearing) -- no atan2 followed by sincos.
__device__ __forceinline__ void kwikoffsets(double lata, double lona, double latb, double lonb, float& dn, float& de, float& ang,
                                            float& dist_nm) {
    dn = (float)((latb - lata) * kDeg2RadD);
    const float dlon = (float)((dmod360((lonb - lona) + 180.0) - 180.0) * kDeg2RadD);
    de = dlon * __cosf((float)((lata + latb) * (0.5 * kDeg2RadD)));
    ang = sqrtf(dn * dn + de * de);
    dist_nm = (kRearth / kNm) * ang;
}

// sin / cos of a latitude in degrees (|lat| <= 90): MUFU.SIN / MUFU.COS, absolute error ~4e-7 on factors of order one
__device__ __forceinline__ void sincos_lat(float latd, float& s, float& c) { __sincosf(latd * kDeg2Rad, &s, &c); }
// sin / cos of an angle in degrees anywhere in (-540, 540) -- a difference of two headings / bearings: folded into
// [-180, 180], where MUFU.SIN / MUFU.COS are good to ~4e-7 absolute (the observations they feed are compared at 5e-4)
__device__ __forceinline__ void sincos_deg(float deg, float& s, float& c) {
    deg = deg > 180.0f ? deg - 360.0f : (deg < -180.0f ? deg + 360.0f : deg);
    __sincosf(deg * kDeg2Rad, &s, &c);
}
// sin of a small angle (differences of nearby positions, a few hundredths of a radian): odd polynomial to x^9, relative
// error < 1e-7 up to |x| = 0.5 -- the MUFU's ABSOLUTE error would be a relative one of 1e-3 at 2 km from a waypoint
__device__ __forceinline__ float sin_small(float x) {
    if (fabsf(x) > 0.5f) return sinf(x);
    const float x2 = x * x;
    float p = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
    p = fmaf(p, x2, 8.3333333e-3f);
    p = fmaf(p, x2, -1.6666667e-1f);
    return fmaf(x * x2, p, x);
}

// geo.rwgs84
__device__ __forceinline__ float rwgs84(float latd) {
    float s, c;
    sincos_lat(latd, s, c);
    const float a = 6378137.0f, b = 6356752.314245f;
    float an = a * a * c, bn = b * b * s, ad = a * c, bd = b * s;
    return sqrtf((an * an + bn * bn) / (ad * ad + bd * bd));
}

// geo.qdrdist (WGS-84 radius haversine).  The bearing's second atan2 argument is rewritten as
// sin(dlat) + 2 sin(lat1) cos(lat2) sin^2(dlon/2), algebraically identical and free of the float32
// cancellation of cos(lat1) sin(lat2) - sin(lat1) cos(lat2) cos(dlon).  qdr in [-180, 180], dist [m].
// `want_dist` (warp-uniform at the call sites): the distance -- WGS-84 radius, two square roots and an atan2 -- is
// only looked at when the FMS timer fires; the bearing is needed every substep.
__device__ __forceinline__ void qdrdist_wgs(double lat1d, double lon1d, double lat2d, double lon2d,
                                            float& qdr, float& dist_m, bool want_dist = true) {
    float la1 = (float)lat1d, la2 = (float)lat2d;
    float dlat = (float)((lat2d - lat1d) * kDeg2RadD);
    float dlon = (float)((lon2d - lon1d) * kDeg2RadD);
    float s1, c1, s2, c2;
    sincos_lat(la1, s1, c1);
    sincos_lat(la2, s2, c2);
    float sh1 = sin_small(0.5f * dlat), sh2 = sin_small(0.5f * dlon);
    dist_m = 0.0f;
    if (want_dist) {
        float r;
        if (la1 * la2 >= 0.0f) {
            r = rwgs84(0.5f * (la1 + la2));
        } else {
            const float a = 6378137.0f;
            r = 0.5f * (fabsf(la1) * (rwgs84(la1) + a) + fabsf(la2) * (rwgs84(la2) + a)) /
                fmaxf(0.000001f, fabsf(la1) + fabsf(la2));
        }
        float root = sh1 * sh1 + c1 * c2 * sh2 * sh2;
        dist_m = 2.0f * r * atan2f(sqrtf(root), sqrtf(fmaxf(0.0f, 1.0f - root)));
    }
    float sdlon = sin_small(dlon);
    qdr = kRad2Deg * atan2f(sdlon * c2, sin_small(dlat) + 2.0f * s1 * c2 * sh2 * sh2);
}

// sub-warp group helpers: G lanes (1, 8, 16 or 32) cooperate on one env.
// Correct under divergence between the groups of one warp: masks name only the caller's group.
template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G >= 32) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31u;
    unsigned base = lane & ~(unsigned)(G - 1);
    return (G == 1) ? (1u << lane) : (((1u << (G & 31)) - 1u) << base);
}
template <int G, typename T> __device__ __forceinline__ T group_bcast(T v, int src) {
    if (G == 1) return v;
    return __shfl_sync(group_mask<G>(), v, src, G);
}
template <int G> __device__ __forceinline__ bool group_all(bool v) {
    if (G == 1) return v;
    if (G >= 32) return __all_sync(0xffffffffu, v);
    const unsigned m = group_mask<G>();
    return (__ballot_sync(m, v) & m) == m;
}
template <int G> __device__ __forceinline__ int group_sum(int v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(group_mask<G>(), v, o, G);
    return v;
}
template <int G> __device__ __forceinline__ unsigned long long group_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(group_mask<G>(), v, o, G);
        v = w < v ? w : v;
    }
    return v;
}

}  // namespace bsg
