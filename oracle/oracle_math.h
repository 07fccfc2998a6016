// TEST INFRASTRUCTURE — CPU restatement oracle of alitteneker/jsraytracer's
// render path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
// PARITY PINNED TO THE REFERENCE ITSELF: the reference ships no golden vectors and the image has
// no JavaScript engine, so the repo brings one (oracle/jsvm) that executes the reference's unmodified src/*.js and
// tests/*/test.mjs; this restatement reproduces those runs BIT FOR BIT — every f32 colour, every ImageData byte — on 34
// of the reference's 37 demo scenes, 31 as committed fixtures (analytic, BVH meshes with vertex normals, SDF, Fresnel, path
// tracing, area lights, depth of field, textures): tests/test_refjs_pin.py, fixtures tests/golden/refjs_*.npz, generator oracle/refjs_golden.py.
// Also pinned by the reference's committed tests/tie_fighter screenshots (tests/test_reference_screenshot.py) and by
// formula-level known-answer tests (tests/test_oracle_kat.py).  Run once, too large to keep as fixtures: dragon / x-wing /
// starwars (profiles/r2_refjs_*_oneoff.log); the three scenes left cannot run from the reference tree.  Caveat: sin / cos / pow are glibc's on both sides.
//
// Numeric model (reference src/math.js:160 `class Vec extends Float32Array`,
// :303 `class Mat extends Array`): vectors are f32 storage, every scalar and
// every matrix entry is f64, every Vec-returning op rounds to f32 on store.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>

namespace orc {

static const double INF = std::numeric_limits<double>::infinity();

struct Vec {
    float v[4];
    int n;
    Vec() : n(0) { v[0] = v[1] = v[2] = v[3] = 0; }
    static Vec of(double a, double b) { Vec r; r.n = 2; r.v[0] = (float)a; r.v[1] = (float)b; return r; }
    static Vec of(double a, double b, double c) { Vec r; r.n = 3; r.v[0] = (float)a; r.v[1] = (float)b; r.v[2] = (float)c; return r; }
    static Vec of(double a, double b, double c, double d) { Vec r; r.n = 4; r.v[0] = (float)a; r.v[1] = (float)b; r.v[2] = (float)c; r.v[3] = (float)d; return r; }
    // Vec.axis(axis, dim, amt) src/math.js:170-174
    static Vec axis(int ax, int dim, double amt = 1) { Vec r; r.n = dim; r.v[ax] = (float)amt; return r; }
    double operator[](int i) const { return (double)v[i]; }
    // JS typed-array read past the end is `undefined`; in arithmetic that is NaN.
    double at(int i) const { return i < n ? (double)v[i] : std::nan(""); }

    // src/math.js:197-214 — map over *this*; b may be longer or shorter.
    Vec plus(const Vec& b) const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = (float)((double)v[i] + b.at(i)); return r; }
    Vec minus(const Vec& b) const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = (float)((double)v[i] - b.at(i)); return r; }
    Vec times(double s) const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = (float)((double)v[i] * s); return r; }
    Vec times(const Vec& b) const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = (float)((double)v[i] * b.at(i)); return r; }
    Vec mult_pairs(const Vec& b) const { return times(b); }
    Vec abs() const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = std::fabs(v[i]); return r; }
    Vec mix(const Vec& b, double s) const { Vec r; r.n = n; for (int i = 0; i < n; ++i) r.v[i] = (float)((1 - s) * (double)v[i] + s * b.at(i)); return r; }
    // src/math.js:252-260 — f64 accumulation left to right over this.length comps
    double dot(const Vec& b) const {
        if (n == 3) return (double)v[0] * b.at(0) + (double)v[1] * b.at(1) + (double)v[2] * b.at(2);
        if (n == 4) return (double)v[0] * b.at(0) + (double)v[1] * b.at(1) + (double)v[2] * b.at(2) + (double)v[3] * b.at(3);
        return (double)v[0] * b.at(0) + (double)v[1] * b.at(1);
    }
    double squarednorm() const { return dot(*this); }
    double norm() const { return std::sqrt(dot(*this)); }
    // src/math.js:242-245
    Vec normalized() const { double nn = norm(); return (nn > 0.00001) ? times(1 / nn) : *this; }
    double sum() const { double a = 0; for (int i = 0; i < n; ++i) a += (double)v[i]; return a; }
    double average() const { return n ? sum() / n : 0; }
    // src/math.js:271-276 (`this[1] || 0`: NaN and undefined become 0)
    static float or0(const Vec& a, int i) { if (i >= a.n) return 0; float x = a.v[i]; return (x != x || x == 0) ? 0.0f : x; }
    Vec to3() const { Vec r; r.n = 3; r.v[0] = v[0]; r.v[1] = or0(*this, 1); r.v[2] = or0(*this, 2); return r; }
    Vec to4(bool isPoint) const { Vec r; r.n = 4; r.v[0] = v[0]; r.v[1] = or0(*this, 1); r.v[2] = or0(*this, 2); r.v[3] = isPoint ? 1.0f : 0.0f; return r; }
    Vec cross(const Vec& b) const {
        return Vec::of((double)v[1] * b[2] - (double)v[2] * b[1], (double)v[2] * b[0] - (double)v[0] * b[2],
                       (double)v[0] * b[1] - (double)v[1] * b[0]);
    }
    static Vec maxs(const Vec& a, double s) { Vec r; r.n = a.n; for (int i = 0; i < a.n; ++i) { double x = a.v[i]; r.v[i] = (float)((x != x || s != s) ? std::nan("") : (x > s ? x : s)); } return r; }
};

struct Mat4 {
    double m[4][4];
    static Mat4 identity() { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = (i == j); return r; }
    // Mat.times(Vec) src/math.js:392-397: result has this.length (4) entries,
    // only the first b.length are written; each is b.dot(row) -> f32.
    Vec times(const Vec& b) const {
        Vec r; r.n = 4;
        for (int i = 0; i < b.n && i < 4; ++i) r.v[i] = (float)dotrow(b, i);
        return r;
    }
    double dotrow(const Vec& b, int r) const {
        if (b.n == 3) return (double)b.v[0] * m[r][0] + (double)b.v[1] * m[r][1] + (double)b.v[2] * m[r][2];
        if (b.n == 4) return (double)b.v[0] * m[r][0] + (double)b.v[1] * m[r][1] + (double)b.v[2] * m[r][2] + (double)b.v[3] * m[r][3];
        return (double)b.v[0] * m[r][0] + (double)b.v[1] * m[r][1];
    }
    // Mat.times(Mat) src/math.js:400-409 (f64, accumulate in index order)
    Mat4 times(const Mat4& b) const {
        Mat4 r;
        for (int i = 0; i < 4; ++i) for (int c = 0; c < 4; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += m[i][k] * b.m[k][c]; r.m[i][c] = s; }
        return r;
    }
    Mat4 transposed() const { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i]; return r; }
    Vec column(int c) const { return Vec::of(m[0][c], m[1][c], m[2][c], m[3][c]); }
};

struct Ray {
    Vec origin, direction;
    Ray() {}
    Ray(const Vec& o, const Vec& d) : origin(o), direction(d) {}
    // src/math.js:294-299
    Ray getTransformed(const Mat4& m) const { return Ray(m.times(origin), m.times(direction)); }
    Vec getPoint(double t) const { return origin.plus(direction.times(t)); }
};

// Math.fmod src/math.js:27: Number((a - floor(a/b)*b).toPrecision(8))
// Number.prototype.toPrecision rounds the EXACT decimal expansion of the double to 8 significant digits and, when two
// 8-digit numbers are equally near, "picks the larger n" (ECMA-262 21.1.3.5 step 10.a; V8's bignum dtoa does the same:
// the last digit is bumped when twice the remainder >= the denominator).  printf rounds such exact ties to even, so they
// are detected and rounded up here.  Ties are not exotic on this path: a coordinate that came from an f32 has few
// significant bits (4.05078125 -> "4.0507813", not "4.0507812"); found by the pin against the reference's own output
// (tests/test_refjs_pin.py, SDF_SphereRepetition).
inline double js_toPrecision8(double x) {
    if (!(x == x) || std::isinf(x) || x == 0) return x;
    char buf[64];
    snprintf(buf, sizeof buf, "%.17e", x);             // [-]d.ddddddd|dddddddddd e+XX (itself correctly rounded)
    const char* p = buf + (buf[0] == '-' ? 1 : 0);
    bool maybe_tie = p[9] == '5';
    for (int i = 10; maybe_tie && i <= 18; ++i) maybe_tie = p[i] == '0';
    if (maybe_tie) {
        char big[900];
        snprintf(big, sizeof big, "%.800e", x);        // every digit a double can have (<= 767 significant)
        const char* q = big + (big[0] == '-' ? 1 : 0);
        bool tie = q[9] == '5';
        const char* e = q + 10;
        for (; tie && *e != 'e'; ++e) tie = *e == '0';
        if (tie) {
            unsigned long long n = 0;
            for (int i = 0; i <= 8; ++i) if (q[i] != '.') n = n * 10 + (unsigned long long)(q[i] - '0');
            const int ex = atoi(strchr(q, 'e') + 1);
            snprintf(buf, sizeof buf, "%s%llue%d", big[0] == '-' ? "-" : "", n + 1, ex - 7);
            return strtod(buf, nullptr);
        }
    }
    snprintf(buf, sizeof buf, "%.7e", x);
    return strtod(buf, nullptr);
}
inline double js_fmod(double a, double b) { return js_toPrecision8(a - (std::floor(a / b) * b)); }
inline double js_sign(double x) { return (x > 0) ? 1.0 : (x < 0 ? -1.0 : x); }
inline double js_max(double a, double b) { if (a != a || b != b) return std::nan(""); return a > b ? a : b; }
inline double js_min(double a, double b) { if (a != a || b != b) return std::nan(""); return a < b ? a : b; }

// ---------------------------------------------------------------------------
// Counter-based RNG shared bit-for-bit with the CUDA kernels
// (jsraytracer_b200/csrc/rng.h).  The reference uses unseeded Math.random()
// (src/renderers.js:95-96 etc.), so only statistical agreement with it is
// possible; keying by (seed, pixel, pass, path-tree node, dimension) lets the
// oracle and the GPU be compared per sample.  SURVEY.md Appendix C.
inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
inline uint32_t rng_sample_key(uint64_t seed, uint32_t pixel, uint32_t pass) {
    uint32_t k = hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + 0x68bc21ebU));
    k = hash32(k + pixel);
    k = hash32(k ^ hash32(pass + 0x9e3779b9U));
    return k;
}
inline uint32_t rng_node_key(uint32_t sample_key, uint32_t node) { return hash32(sample_key + 0x85ebca6bU * node); }
// path-tree node id of child `which` (0 reflection, 1 transmission / refraction): 2k + which while that fits
// (levels <= 30), a hash with the top bit set beyond (never 0 or 1, never back in the doubling range)
inline uint32_t rng_child_node(uint32_t node, uint32_t which) {
    return (node < 0x40000000U) ? 2U * node + which : (hash32(node ^ (0x9e3779b9U + which)) | 0x80000000U);
}
inline double rng_u01(uint32_t node_key, uint32_t dim) {
    return (double)(hash32(node_key + 0xc2b2ae35U * (dim + 1)) >> 8) * (1.0 / 16777216.0);
}
// tape mode: one sequential splitmix64 stream per pixel sample (the same few lines in Python: oracle/refjs.py)
inline uint64_t rng_tape_seed(uint64_t seed, uint32_t pixel, uint32_t pass) { return (seed << 48) ^ ((uint64_t)pass << 32) ^ (uint64_t)pixel; }
inline double rng_tape_next(uint64_t& state) {
    state += 0x9E3779B97F4A7C15ULL;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
enum { DIM_JITTER_X = 0, DIM_JITTER_Y = 1, DIM_LENS_A = 2, DIM_LENS_R = 3, DIM_LIGHTS = 8 };

}  // namespace orc
