// NOT a scene of the reference: aggregates nested the ways the reference's API allows (src/aggregates.js:14-18,43-49 —
// `ancestors.unshift(this)` at every level) and none of its own scenes exercises: a BVHAggregate as a member of a plain,
// transformed Aggregate (next to Primitives), and a BVHAggregate built over two more instances of that tree plus a Primitive
// (BVHAggregate.build takes any WorldObject with a finite bounding box, src/aggregates.js:34-42).  The JavaScript twin of
// jsraytracer_b200/scenes nested_aggregates(); run by oracle/refjs.py like any other test.
export function configureTest(callback) {

    const camera = new PerspectiveCamera(Math.PI / 4, 1,
        Mat4.translation([0, 1.2, 6]).times(Mat4.rotation(-0.15, Vec.of(1,0,0))));
    const lights = [new SimplePointLight(Vec.of(8, 9, 10, 1), Vec.of(1, 1, 1), 6000)];

    const objs = [new Primitive(
        new Plane(),
        new PhongMaterial(new CheckerboardMaterialColor(Vec.of(1,1,1), Vec.of(0.1,0.1,0.1)), 0.2, 0.4, 0.6, 100, 0.4),
        Mat4.translation([0,-1,0]).times(Mat4.rotation(Math.PI/2, Vec.of(1,0,0))))];

    loadObjFile(
        "../assets/tetrahedron.obj",
        new PhongMaterial(Vec.of(1, 0.3, 0.2), 0.2, 0.5, 0.5, 50, 0.3),

        function(tris) {
            const mesh = BVHAggregate.build(tris,
                Mat4.translation([-1.6, 0, -4]).times(Mat4.rotation(0.5, Vec.of(0,1,0))));
            const ball = new Primitive(new Sphere(),
                new PhongMaterial(Vec.of(0.2, 0.3, 1), 0.2, 0.4, 0.6, 100, 0.5),
                Mat4.translation([1.4, 0.2, -1.5]).times(Mat4.scale(0.6)));
            objs.push(new Aggregate(
                [ball, mesh, new Primitive(new UnitBox(),
                    new PhongMaterial(Vec.of(0.2, 0.9, 0.3), 0.2, 0.4, 0.6, 100, 0.3),
                    Mat4.translation([0, -0.5, -6]))],
                Mat4.translation([0.3, 0, 0]).times(Mat4.rotation(0.2, Vec.of(0,1,0)))));

            const inst = [-2.5, 2.2].map(x => new BVHAggregate(tris, mesh.kdtree,
                Mat4.translation([x, 0.1, -7]).times(Mat4.rotation(0.4 * x, Vec.of(0,1,0))).times(Mat4.scale(1.3))));
            const moon = new Primitive(new Sphere(),
                new FresnelPhongMaterial(Vec.of(1, 1, 0.4), 0.1, 0.4, 0.5, 100, 1.4),
                Mat4.translation([0, 2.4, -7]).times(Mat4.scale(0.8)));
            objs.push(BVHAggregate.build(inst.concat([moon]), Mat4.translation([0, 0.2, 0])));

            callback({
                renderer: new IncrementalMultisamplingRenderer(
                    new World(objs, lights, Vec.of(0.1, 0.1, 0.15)), camera, 16, 4),
                width: 600,
                height: 600
            });
        });
}
