// NOT a scene of the reference: written for this repository in the shape of the reference's tests/*/test.mjs, out of the
// reference's own classes, to reach what none of its 37 demo scenes uses — TransparentMaterial, SolidColorMaterial,
// PositionalUVMaterial, RoundSDF, a Circle area light, TextureMaterialColor in "nearest" mode and with wrapped
// coordinates, a texture under a ScaledMaterialColor, nested checkerboards, Cylinder under a Fresnel material, a background colour.
// Run by oracle/refjs.py like any other test (fixture tests/golden/refjs_extra_materials.npz).
export function configureTest(callback) {

    const camera = new PerspectiveCamera(Math.PI / 4, 1,
        Mat4.translation([0, 1.2, 2]).times(Mat4.rotation(-0.25, Vec.of(1,0,0))));

    const img = new ImageData(8, 4);
    for (let y = 0; y < 4; ++y)
        for (let x = 0; x < 8; ++x) {
            const i = 4 * (y * 8 + x);
            img.data[i] = 32 * x + 15;
            img.data[i + 1] = 60 * y + 20;
            img.data[i + 2] = (x * y * 9) % 256;
            img.data[i + 3] = 255;
        }
    const texBilinearWrap = new TextureMaterialColor(img, "bilinear", false, false);
    const texNearestClamp = new TextureMaterialColor(img, "nearest", true, true);

    const lights = [
        new SimplePointLight(Vec.of(3, 6, 4, 1), Vec.of(1, 0.9, 0.8), 900),
        new RandomSampleAreaLight(new Circle(),
            Mat4.translation([-2, 4, -2]).times(Mat4.rotation(Math.PI/2, Vec.of(1,0,0))),
            Vec.of(0.6, 0.7, 1), 400, 2)];

    const objects = [];
    // floor: positional UVs into a wrapped bilinear texture, through Phong with a mirror term
    objects.push(new Primitive(
        new Plane(),
        new PositionalUVMaterial(
            new PhongMaterial(Vec.of(1,1,1), 0.2, texBilinearWrap, 0.3, 20, 0.2),
            Vec.of(0.3, 0, 0.1, 1), Vec.of(0.37, 0, 0, 0), Vec.of(0, 0, 0.41, 0)),
        Mat4.translation([0,-1,0]).times(Mat4.rotation(Math.PI/2, Vec.of(1,0,0)))));
    // a pane of coloured glass that casts no shadow
    objects.push(new Primitive(
        new Square(),
        new TransparentMaterial(Vec.of(0.9, 0.2, 0.2), 0.35),
        Mat4.translation([-0.8, 0, -3]).times(Mat4.scale([2, 2, 1])),
        undefined, false));
    // an unlit ball: SolidColorMaterial over the nearest / clamped texture through the sphere's UVs
    objects.push(new Primitive(
        new Sphere(),
        new SolidColorMaterial(texNearestClamp),
        Mat4.translation([1.2, 0, -4])));
    // a square whose colour is a checkerboard of a scaled solid and another checkerboard
    objects.push(new Primitive(
        new Square(),
        new PhongMaterial(
            new CheckerboardMaterialColor(new ScaledMaterialColor(Vec.of(1, 0.5, 0.25), 0.8),
                new CheckerboardMaterialColor(Vec.of(0.1, 0.6, 0.2), Vec.of(0.9, 0.9, 0.1))),
            0.2, 0.5, 0.4, 30, 0.1),
        Mat4.translation([-1.8, 0.2, -5]).times(Mat4.rotation(0.5, Vec.of(0,1,0))).times(Mat4.scale([3, 3, 1]))));
    // a square whose base colour is the texture scaled per channel
    objects.push(new Primitive(
        new Square(),
        new PhongMaterial(new ScaledMaterialColor(texBilinearWrap, 0.8), 0.3, 0.6, 0.2, 10),
        Mat4.translation([0.9, 1.6, -5.5]).times(Mat4.scale([1.6, 1.2, 1]))));
    // a rounded box as an SDF
    objects.push(new Primitive(
        new SDFGeometry(new RoundSDF(new BoxSDF(Vec.of(0.5, 0.3, 0.4)), 0.15)),
        new PhongMaterial(Vec.of(0.2, 0.4, 1), 0.2, 0.5, 0.5, 40, 0.3),
        Mat4.translation([0.2, -0.4, -2.5]).times(Mat4.rotation(0.7, Vec.of(0,1,0)))));
    // a glass cylinder
    objects.push(new Primitive(
        new Cylinder(),
        new FresnelPhongMaterial(Vec.of(1,1,1), 0.05, 0.2, 0.6, 60, 1.4),
        Mat4.translation([2.2, -0.3, -6])));

    callback({
        renderer: new IncrementalMultisamplingRenderer(
            new World(objects, lights, Vec.of(0.05, 0.07, 0.1)), camera, 4, 5),
        width: 300,
        height: 200
    });
}
