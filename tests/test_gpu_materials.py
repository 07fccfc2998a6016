"""Materials / geometry of SURVEY §8f item 3 that no BASELINE config uses: SolidColorMaterial,
TransparentMaterial, a non-unit AABB geometry, RoundSDF, does_cast_shadow = false."""
import numpy as np
import pytest

from conftest import psnr

pytestmark = pytest.mark.gpu


def _scene():
    from jsraytracer_b200.jsmath import Vec, Mat4
    from jsraytracer_b200.geometry import AABB, Plane, Sphere, UnitBox
    from jsraytracer_b200.materials import (CheckerboardMaterialColor, PhongMaterial, SolidColorMaterial,
                                             TransparentMaterial, SimplePointLight)
    from jsraytracer_b200.sdf import SDFGeometry, RoundSDF, BoxSDF
    from jsraytracer_b200.world import World, Primitive
    from jsraytracer_b200.cameras import PerspectiveCamera
    from jsraytracer_b200.renderers import SimpleRenderer
    import math
    cam = PerspectiveCamera(math.pi / 4, 1, Mat4.translation([0, 1.5, 6]).times(Mat4.rotation(-0.15, Vec.of(1, 0, 0))))
    lights = [SimplePointLight(Vec.of(6, 8, 8, 1), Vec.of(1, 1, 1), 3000)]
    objs = [
        Primitive(Plane(), PhongMaterial(CheckerboardMaterialColor(Vec.of(1, 1, 1), Vec.of(0.2, 0.2, 0.2)), 0.1, 0.6, 0.3, 20, 0.2),
                  Mat4.translation([0, -1, 0]).times(Mat4.rotation(math.pi / 2, Vec.of(1, 0, 0)))),
        Primitive(Sphere(), SolidColorMaterial(Vec.of(0.9, 0.2, 0.1)), Mat4.translation([-2.2, 0, -2])),
        Primitive(UnitBox(), TransparentMaterial(Vec.of(0.1, 0.4, 0.9), 0.35), Mat4.translation([0, 0, -1]).times(Mat4.scale(1.6))),
        Primitive(AABB(Vec.of(0.25, 0, 0, 1), Vec.of(0.5, 1.0, 0.25, 0)), PhongMaterial(Vec.of(0.2, 0.8, 0.3), 0.2, 0.5, 0.4, 30, 0.3),
                  Mat4.translation([2.2, 0, -2.5]).times(Mat4.rotation(0.5, Vec.of(0, 1, 0)))),
        Primitive(SDFGeometry(RoundSDF(BoxSDF(Vec.of(0.4, 0.3, 0.4), Vec.of(0.9, 0.8, 0.1)), 0.15), 200, 0.0001, 100),
                  PhongMaterial(Vec.of(1, 1, 1), 0.2, 0.6, 0.3, 40, 0.2), Mat4.translation([0.2, 1.6, -3])),
        Primitive(Sphere(), PhongMaterial(Vec.of(1, 1, 1), 0.5, 0.4, 0.2, 10, 0), Mat4.translation([1.5, 2.5, 1]).times(Mat4.scale(0.5)),
                  None, False),        # does not cast a shadow (src/world.js:117-118)
    ]
    return {"renderer": SimpleRenderer(World(objs, lights, Vec.of(0.1, 0.1, 0.15)), cam, 4), "width": 256, "height": 256}


def test_misc_materials_and_geometry_match_oracle():
    from jsraytracer_b200 import lib
    from jsraytracer_b200.serializer import Serializer
    from oracle.oracle import OracleScene
    ser = Serializer(_scene())
    sc = lib.Scene(ser.to_msgpack(), lib.FORMAT_MSGPACK, device=0)
    orc = OracleScene(ser.to_json())
    ids, t = sc.primary_hits()
    oids, ot, _ = orc.primary_hits()
    assert float((ids == oids).mean()) >= 0.9999
    assert set(np.unique(oids).tolist()) >= {0, 1, 2, 3, 4, 5}           # every primitive is visible
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    g = np.clip(sc.read_accum()[0][..., :3], 0, 1)
    o = np.clip(orc.render(1, seed=1, jitter=False)[0], 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
