// Node harness for the reference's CPU path and the CUDA path side by side (cannot run in the
// build image: no node).  Modelled on tests/test_to_json.js:1-22: the reference's sources are
// browser scripts, so they are loaded with vm.runInThisContext, with shims for the browser
// globals they touch (ImageData for src/pixelbuffer.js:1-8, fspromise for src/objloader.js:22-25).
//
//   node js/worker_harness.js <reference root> <test name> [cuda|cpu] [workers]
//
// cpu:  worker_threads, one per core, columns interleaved exactly like src/worker.js:30-32
//       (renderer.render(buffer, 1000, cb, workerIndex, workerCount)), composited by alpha like
//       src/raytrace_launcher.js:92-97.
// cuda: a CUDARenderer built from the same world/camera (js/cuda_renderer.js).
"use strict";
const fs = require("fs"), path = require("path"), vm = require("vm"), os = require("os");
const { Worker, isMainThread, parentPort, workerData } = require("worker_threads");

function loadReference(root) {
    global.fs = fs; global.fspromise = require("fs/promises");
    global.ImageData = class ImageData { constructor(w, h) { this.width = w; this.height = h; this.data = new Uint8ClampedArray(w * h * 4); } };
    for (const f of ["math.js", "world.js", "pixelbuffer.js", "geometry.js", "materials.js", "cameras.js", "renderers.js",
                     "lights.js", "objloader.js", "sdf.js", "aggregates.js", "serializer.js"])
        new vm.Script(fs.readFileSync(path.join(root, "src", f), "utf8"), { filename: f }).runInThisContext();
}

async function configure(root, name) {
    const mod = await import(path.join(root, "tests", name, "test.mjs"));
    process.chdir(path.join(root, "tests"));        // the tests load "../assets/..."
    return new Promise(resolve => mod.configureTest(resolve));
}

async function main() {
    const [root, name, mode = "cuda", nworkers = os.cpus().length] = process.argv.slice(2);
    if (!isMainThread) {
        loadReference(workerData.root);
        const test = await configure(workerData.root, workerData.name);
        const buffer = new PixelBuffer(test.width, test.height);
        test.renderer.render(buffer, 0, false, workerData.index, workerData.count);     // src/worker.js:30-32
        parentPort.postMessage(buffer.imgdata.data);
        return;
    }
    loadReference(root);
    const test = await configure(root, name);
    const t0 = Date.now();
    if (mode === "cuda") {
        vm.runInThisContext(fs.readFileSync(path.join(__dirname, "cuda_renderer.js"), "utf8"));
        const r = test.renderer;
        const cuda = new CUDARenderer(r.world, r.camera, r.samplesPerPixel || 1, r.maxRecursionDepth);
        const img = cuda.render(new PixelBuffer(test.width, test.height));
        console.log(`cuda: ${(Date.now() - t0) / 1000}s`, img.imgdata.data.length);
    } else {
        const out = new Uint8ClampedArray(test.width * test.height * 4);
        await Promise.all([...Array(+nworkers).keys()].map(i => new Promise(res => {
            new Worker(__filename, { workerData: { root, name, index: i, count: +nworkers } }).on("message", data => {
                for (let p = 0; p < out.length; p += 4) if (data[p + 3]) out.set(data.subarray(p, p + 4), p);   // alpha compositing of the column stripes
                res();
            });
        })));
        console.log(`cpu x${nworkers}: ${(Date.now() - t0) / 1000}s`);
    }
}
main();
