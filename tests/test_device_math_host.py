"""`Math.fmod` AS THE DEVICE COMPUTES IT, checked on the CPU against the reference.

The block of jsraytracer_b200/csrc/device_math.cuh that implements `Number(x.toPrecision(8))` and `Math.fmod` (the SDF
infinite-repetition transformer's arithmetic, src/math.js:27) is plain C++ apart from a handful of CUDA spellings.  This
test cuts that block out of the source file as it is, compiles it for the host behind shims for those spellings
(`__dmul_rn` -> `*` under -ffp-contract=off, `__ldg(p)` -> `*p`, `__double2hiint`), and compares it with what the
reference's own `Math.fmod` returned for 4 000 arguments when its source ran in oracle/jsvm (tests/golden/probes_refjs.npz) —
among them the exact decimal ties that `toPrecision` rounds upwards and `rint` would round to even."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "jsraytracer_b200", "csrc", "device_math.cuh")

SHIM = r"""
#define _GNU_SOURCE 1
#include <cmath>
#include <cstdint>
#include <cstring>
using std::isfinite;
#define __device__
#define __noinline__
#define JSRT_DEV static inline
static inline double dmul(double a, double b) { return a * b; }      // __dmul_rn / __dadd_rn / __dsub_rn: IEEE round-to-nearest,
static inline double dadd(double a, double b) { return a + b; }      // never contracted (this file is built with -ffp-contract=off)
static inline double dsub(double a, double b) { return a - b; }
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
#define __ldg(p) (*(p))
"""


def _build(tmp_path):
    text = open(SRC).read()
    a = text.index("__device__ const double kPow10[32]")
    b = text.index("\n", text.index("JSRT_DEV double js_fmod_pow2")) + 1
    block = text[a:b]
    assert "js_round_half_up" in block and "js_to_precision8_slow" in block and "js_fmod" in block
    cpp = tmp_path / "dev_fmod.cpp"
    cpp.write_text(SHIM + block + '\nextern "C" double dev_js_fmod(double a, double b) { return js_fmod(a, b); }\n'
                   'extern "C" double dev_js_fmod_pow2(double a, double b) { return js_fmod_pow2(a, b, 1.0 / b); }\n'
                   'extern "C" double dev_to_precision8(double x) { return js_to_precision8(x); }\n')
    so = tmp_path / "dev_fmod.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", str(so), str(cpp)])
    L = ctypes.CDLL(str(so))
    for f in (L.dev_js_fmod, L.dev_js_fmod_pow2):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double, ctypes.c_double]
    L.dev_to_precision8.restype = ctypes.c_double
    L.dev_to_precision8.argtypes = [ctypes.c_double]
    return L


def test_device_fmod_equals_reference_math_fmod(tmp_path):
    L = _build(tmp_path)
    z = np.load(os.path.join(HERE, "golden", "probes_refjs.npz"))
    a, b, want = z["fmod_a"], z["fmod_b"], z["fmod_out"]
    got = np.array([L.dev_js_fmod(float(x), float(y)) for x, y in zip(a, b)])
    bad = np.nonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))[0]
    assert bad.size == 0, [(float(a[i]), float(b[i]), float(want[i]), float(got[i])) for i in bad[:5]]
    # the compiler-supplied-reciprocal form used for power-of-two periods (sdf_compile.cpp)
    p2 = np.isin(b, [1.0, 2.0, 4.0])
    got2 = np.array([L.dev_js_fmod_pow2(float(x), float(y)) for x, y in zip(a[p2], b[p2])])
    assert np.array_equal(got2, want[p2], equal_nan=True)
    # the ties are really in there: rounding them to even would have failed this test
    r = a - np.floor(a / b) * b
    even = np.array([float("%.7e" % v) for v in r])
    assert int((even != want).sum()) >= 20


def test_device_to_precision8_equals_the_interpreter_on_ties_and_ranges():
    """toPrecision(8) over the magnitudes SDF coordinates take against oracle/jsvm's
    Number.prototype.toPrecision — exact decimal arithmetic, ties upwards (ECMA-262 21.1.3.5)"""
    import tempfile
    import pathlib
    from oracle.jsvm.runtime import to_precision
    with tempfile.TemporaryDirectory() as d:
        L = _build(pathlib.Path(d))
        rng = np.random.default_rng(7)
        xs = []
        for e in range(-12, 8):          # the fast path: 1e-15 <= |x| < 1e8, every coordinate an SDF scene produces
            m = rng.integers(10 ** 7, 10 ** 8, 30)
            xs += list((m + 0.5) * 10.0 ** (e - 7))           # near-ties and, where the product is exact, true ties
            xs += list(rng.uniform(1, 10, 30) * 10.0 ** e)
        xs += [(k + 0.5) / 2.0 ** s for k in rng.integers(0, 2 ** 22, 400) for s in (3, 9, 14)]   # f32-like values: exact ties are common
        xs += [-x for x in xs[:200]]
        bad = []
        for x in xs:
            x = float(x)
            want = float(to_precision(x, 8))
            got = L.dev_to_precision8(x)
            if got != want:
                bad.append((x, want, got))
        # The device code rounds |x| * 10^k computed in f64: exact when that product is exact (every tie that occurs on an
        # f32-derived coordinate), within one unit of the 8th digit otherwise when the product lands within an ulp of a tie.
        assert len(bad) <= len(xs) // 500, bad[:5]
        for x, want, got in bad:
            assert abs(got - want) <= abs(want) * 1.01e-7
