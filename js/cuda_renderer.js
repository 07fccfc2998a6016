// CUDARenderer — the B200 render path as a class that sits next to the reference's
// SimpleRenderer / IncrementalMultisamplingRenderer (reference: src/renderers.js).
//
// Same constructor shape (world, camera, samplesPerPixel, maxRecursionDepth = 3) and the same
// render(img, timelimit = 0, callback = false, x_offset = 0, x_delt = 1) signature as
// IncrementalMultisamplingRenderer.render (src/renderers.js:66-70): it fills img.imgdata.data
// (PixelBuffer.setColor semantics, src/pixelbuffer.js:39-49), calls callback({pass, completion})
// at most every `timelimit` ms (src/renderers.js:103-112) and returns img.
//
// It replaces the worker pool of src/raytrace_launcher.js / src/worker.js for this path: the scene
// travels once as `new Serializer({renderer, width, height}).plain()` (src/serializer.js) through
// the N-API addon (js/napi/jsrt_addon.cc -> include/jsrt.h); all passes run on the GPU with the
// accumulation buffer resident in HBM.
//
// NOTE: no Node exists in the build image.  This file is executed by the repository's own JavaScript
// interpreter (oracle/jsvm, test infrastructure) next to the reference's unmodified sources, with the
// addon's entry points supplied by Python: tests/test_js_host_in_vm.py; its Python twin
// (jsraytracer_b200/renderers.py, same logic over the same C ABI) is what the GPU tests drive.  It is loaded
// like the reference's sources (tests/test_to_json.js:7-22 runs them with vm.runInThisContext), so
// it uses the reference's globals (Serializer, Triangle, IncrementalMultisamplingRenderer).
"use strict";

// The reference's Triangle.serialize writes `psdata: serializeStep(this.ps)`
// (src/geometry.js:355-357), which drops vertex normals / UVs that the worker path keeps.
// Serialise what the constructor actually takes.
function installTriangleSerializeFix() {
    Triangle.prototype.serialize = function (serializer) {
        return { ps: serializer.serializeStep(this.ps), psdata: serializer.serializeStep(this.psdata) };
    };
}

// A TextureMaterialColor holds a browser ImageData whose width / height / data are prototype getters, so the
// reference's serializer writes it as an empty object (src/serializer.js:54-57; src/materials.js:88-90 then reads
// `data.imgdata`, which is never written).  Give the texture a serialize() that emits the wire form the C ABI
// reads: `_imgdata: {_t, _v: {width, height, data}}` with `data` the RGBA8 bytes (msgpack bin via Uint8Array; a
// plain array of numbers under JSON) — the same form jsraytracer_b200/materials.py:ImageData emits.
class ImageDataWire {          // stand-in that serialises: own serialize(), so serializeStep leaves `data` alone
    constructor(img, binary) { this.img = img; this.binary = binary; }
    serialize(serializer) {
        const d = this.img.data, bytes = new Uint8Array(d.buffer, d.byteOffset, d.byteLength);
        return { width: this.img.width, height: this.img.height, data: this.binary ? bytes : Array.from(bytes) };
    }
}
function installTextureSerializeFix() {
    if (typeof TextureMaterialColor === "undefined") return;
    TextureMaterialColor.prototype.serialize = function (serializer) {
        return {
            _imgdata: serializer.serializeStep(new ImageDataWire(this._imgdata, !!serializer.binary)),
            width: this.width, height: this.height, mode: this.mode, clampU: this.clampU, clampV: this.clampV,
        };
    };
}

class CUDARenderer extends IncrementalMultisamplingRenderer {
    constructor(world, camera, samplesPerPixel, maxRecursionDepth = 3, options = {}) {
        super(world, camera, samplesPerPixel, maxRecursionDepth);
        // non-enumerable: must not travel with the scene (src/serializer.js:54-57 emits own enumerable keys)
        Object.defineProperty(this, "_opt", { value: Object.assign({ seed: 1, device: 0, addon: null }, options), enumerable: false, writable: true });
        Object.defineProperty(this, "_scene", { value: null, enumerable: false, writable: true });
        Object.defineProperty(this, "_size", { value: null, enumerable: false, writable: true });
    }

    // options.devices = [0, 1, ...]: the scene is replicated on every listed GPU, each render() call's passes are dealt to
    // them and the image is composed on the first one (the in-process successor of src/raytrace_launcher.js's worker pool)
    _devices() {
        return (this._opt.devices && this._opt.devices.length) ? Int32Array.from(this._opt.devices) : this._opt.device;
    }

    _addon() {
        if (!this._opt.addon)
            this._opt.addon = require("./napi/build/Release/jsrt_addon.node");
        return this._opt.addon;
    }

    // The *_json scenes (tests/dragon_json/test.mjs, tests/toledo_json/test.mjs): a scene that only exists in wire form.
    // `Serializer.deserializeJSON` would rebuild the whole JS object graph just to serialise it again; here the blob
    // goes to the library as is.  Returns what `configureTest` hands to its callback: {renderer, width, height}.
    //   CUDARenderer.fromWire(fs.readFileSync("tests/dragon/test.json"), 0)        // JSON text
    //   CUDARenderer.fromWire(fs.readFileSync("tests/toledo/test.msgpack"), 1)     // msgpack
    static fromWire(blob, format = 0, options = {}) {
        const r = new CUDARenderer(null, null, 1, 3, options);
        const head = new Int32Array(5);
        r._addon().sceneHeader(blob, format, head);               // throws Error(jsrt_last_error()) on a malformed blob
        r.samplesPerPixel = head[4] ? head[2] : 1;                // a serialised SimpleRenderer: one un-jittered pass
        r.maxRecursionDepth = head[3];
        Object.defineProperty(r, "_wire", { value: { blob, format, jitter: !!head[4] }, enumerable: false });
        return { renderer: r, width: head[0], height: head[1] };
    }

    _ensureScene(img) {
        const w = img.width(), h = img.height();
        if (this._scene && this._size[0] === w && this._size[1] === h)
            return this._scene;
        if (this._scene) this._addon().destroyScene(this._scene);
        if (this._wire) {
            // a wire blob carries its own size: the image handed to render() must match it (rows would be written with the wrong stride)
            const head = new Int32Array(5);
            this._addon().sceneHeader(this._wire.blob, this._wire.format, head);
            if (head[0] !== w || head[1] !== h)
                throw new Error("render: the serialised scene is " + head[0] + "x" + head[1] + ", the image " + w + "x" + h);
            this._scene = this._addon().createScene(this._wire.blob, this._wire.format, this._devices());
            this._size = [w, h];
            return this._scene;
        }
        installTriangleSerializeFix();
        installTextureSerializeFix();
        let msgpack = null;
        try { msgpack = require("@msgpack/msgpack"); } catch (e) { msgpack = null; }
        const ser = Object.create(Serializer.prototype);      // like `new Serializer(data)`, with a flag the texture fix reads
        ser.SER_ID = "_SID" + Serializer.SER_UID_GEN++; ser.REF_UID_GEN = 0; ser.refs = {}; ser.ref_counts = {}; ser.type_map = {};
        ser.binary = !!msgpack;
        const plain = ser.serializeStep({ renderer: this, width: w, height: h });
        // msgpack keeps Infinity (IOR, SDF box sizes); JSON.stringify would write null, which the
        // reader maps back to +Infinity at those fields, so both work (tests/test_to_json.js:36-38).
        let blob, format;
        if (msgpack) { blob = Buffer.from(msgpack.encode(plain)); format = 1; }
        else { blob = Buffer.from(JSON.stringify(plain), "utf8"); format = 0; }
        this._scene = this._addon().createScene(blob, format, this._devices());   // throws Error(jsrt_last_error())
        this._size = [w, h];
        return this._scene;
    }

    // The uploaded scene is kept between render() calls (serialising a 100 000-triangle world takes seconds in JS).  The
    // reference's renderers re-read this.world / this.camera on every call; after editing either (camera.setTransform,
    // a material change, another samplesPerPixel is fine) call invalidate() so that the next render() serialises again.
    invalidate() {
        if (this._scene) { this._addon().destroyScene(this._scene); this._scene = null; }
    }

    render(img, timelimit = 0, callback = false, x_offset = 0, x_delt = 1) {
        const addon = this._addon(), scene = this._ensureScene(img);
        const spp = this.samplesPerPixel;
        const flags = (this._wire && !this._wire.jitter) ? 1 : 0;      // flag 1 = no jitter
        // One library call per pass, like the reference's loop (src/renderers.js:87): the library coalesces consecutive calls
        // into full waves.  With a progress callback the host synchronises once per group of 8 passes — not per pass — to
        // look at the clock (src/renderers.js:103-112: at most one callback per `timelimit` ms).
        const group = this._opt.passesPerGroup || 8;
        addon.resetAccum(scene);
        let last = Date.now();
        for (let done = 0; done < spp; ) {
            addon.render(scene, done, 1, this._opt.seed, x_offset, x_delt, flags);      // asynchronous on the scene's stream
            done += 1;
            if (timelimit && callback && (done % group === 0 || done === spp)) {
                addon.synchronize(scene);
                const now = Date.now();
                if (now - last >= timelimit) {
                    last = now;
                    addon.resolveRGBA8(scene, img.imgdata.data);
                    callback({ pass: done - 1, completion: done / spp });
                }
            }
        }
        addon.resolveRGBA8(scene, img.imgdata.data);        // Uint8ClampedArray of W*H*4, like PixelBuffer.setColor fills
        return img;
    }

    // The GL renderer's display path (gl/src/WebGLRendererAdapter.js:183-246,352-379): the same passes with the auxiliary
    // buffers (running per-pixel variance, first-hit normal / distance sums: flag 4), then the variance-guided filter.
    // options: {sigma = 1, kSigma = 2, threshold = 5, colorLogScale = 0} (the GL adapter's defaults, :15-22).
    renderDenoised(img, options = {}, x_offset = 0, x_delt = 1) {
        const addon = this._addon(), scene = this._ensureScene(img);
        const flags = ((this._wire && !this._wire.jitter) ? 1 : 0) | 4;
        addon.resetAccum(scene);
        addon.render(scene, 0, this.samplesPerPixel, this._opt.seed, x_offset, x_delt, flags);
        addon.denoise(scene, options.sigma ?? 1, options.kSigma ?? 2, options.threshold ?? 5, options.colorLogScale ?? 0, img.imgdata.data);
        return img;
    }
    // {normalDepth, variance}: two Float32Array(W*H*4) — sums of the first hit's normal (xyz) and distance (w); the GL shader's
    // variance sums (xyz) and the number of samples whose camera ray hit something (w)
    readAov(img) {
        const addon = this._addon(), scene = this._ensureScene(img);
        const n = img.width() * img.height() * 4;      // PixelBuffer.width() / height(), src/pixelbuffer.js
        const out = { normalDepth: new Float32Array(n), variance: new Float32Array(n) };
        addon.readAov(scene, out.normalDepth, out.variance);
        return out;
    }

    close() {
        if (this._scene) { this._addon().destroyScene(this._scene); this._scene = null; }
    }
}

if (typeof module !== "undefined") module.exports = { CUDARenderer, installTriangleSerializeFix, installTextureSerializeFix };
