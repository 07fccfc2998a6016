"""Host mirror of the reference's numeric model (src/math.js)."""
import math

import numpy as np

from jsraytracer_b200.jsmath import Vec, Mat, Mat4, Ray, median, quick_select, fmod, js_to_precision8, f32


def test_vec_ops_round_to_f32():
    a = Vec.of(0.1, 0.2, 0.3)
    assert a[0] == float(np.float32(0.1))                      # Float32Array store
    s = a.plus(Vec.of(1, 1, 1))
    assert s[0] == float(np.float32(np.float64(np.float32(0.1)) + 1.0))
    assert a.dot(a) == a[0] * a[0] + a[1] * a[1] + a[2] * a[2]  # f64 accumulation, not rounded
    # map iterates over `this`: a 3-vector plus a 4-vector stays 3 long (renderers.js:93-97)
    assert len(Vec.of(0, 0, 0).plus(Vec.of(1, 2, 3, 1))) == 3


def test_normalized_keeps_tiny_vectors():
    v = Vec.of(1e-6, 0, 0)
    assert v.normalized() is v                                 # src/math.js:242-245
    n = Vec.of(3, 0, 4).normalized()
    assert abs(n.norm() - 1) < 1e-6


def test_to3_to4():
    assert Vec.of(1, 2).to4(False).tolist() == [1, 2, 0, 0]
    assert Vec.of(1, 2, 3, 9).to3().tolist() == [1, 2, 3]
    assert Vec.of(1, 2, 3).to4(True).tolist() == [1, 2, 3, 1]


def test_mat_times_vec_length_rule():
    m = Mat4.translation([1, 2, 3])
    assert m.times(Vec.of(0, 0, 0, 1)).tolist() == [1, 2, 3, 1]
    # a 3-vector only fills 3 rows and sees no translation (src/math.js:392-397)
    assert m.times(Vec.of(1, 1, 1)).tolist() == [1, 1, 1, 0]


def test_mat4_inverse_roundtrip():
    m = Mat4.translation([1.2, 0.2, -7]).times(Mat4.rotation(0.35, Vec.of(0, 1, 0))).times(Mat4.scale(2.5))
    p = m.times(Mat4.inverse(m))
    for i in range(4):
        for j in range(4):
            assert abs(p[i][j] - (1 if i == j else 0)) < 1e-12


def test_rotation_matches_axis_forms():
    a = 0.7
    rx, r = Mat4.rotationX(a), Mat4.rotation(a, Vec.of(1, 0, 0))
    for i in range(4):
        for j in range(4):
            assert abs(rx[i][j] - r[i][j]) < 1e-15


def test_ray_point_two_roundings():
    r = Ray(Vec.of(0.1, 0, 0, 1), Vec.of(0.3, 0, 0, 0))
    t = 1.7
    expect = f32(f32(0.1) + f32(f32(0.3) * t))
    assert r.getPoint(t)[0] == expect


def test_fmod_to_precision8():
    assert fmod(5.3, 2) == 1.3
    assert fmod(-0.25, 2) == 1.75
    assert js_to_precision8(1 / 3) == 0.33333333
    assert js_to_precision8(123456789.0) == 123456790.0


def test_quickselect_and_median_reference_quirks():
    rng = np.random.default_rng(0)
    for n in (1, 3, 7, 101, 1501):
        a = rng.normal(size=n).tolist()
        assert median(list(a)) == sorted(a)[n // 2]
    # even length: the reference averages elements len/2 and len/2+1 (sic) of the
    # partially ordered array, not the two middle ones
    assert median([1., 2., 3., 4.]) == 3.5
    # length 2 reads one past the end of the JS array; reproduced as in V8
    assert median([3.0, 1.0]) == 2.0 and median([5.0, 5.0]) == 5.0
    a = rng.normal(size=50).tolist()
    assert quick_select(list(a), 10) == sorted(a)[10]
