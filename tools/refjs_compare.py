"""One scene of the reference, run now in oracle/jsvm, against the oracle's tape mode (needs /root/reference):
    python tools/refjs_compare.py <scene> <width> <height> [passes]
prints the interpreter's times, the Math.random() draws per sample, and how many pixels differ (f32 colours, ImageData bytes).
The committed fixtures (oracle/refjs_golden.py) are the same comparison frozen; this is for scenes too large to keep."""
import sys, time
import numpy as np
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from oracle.refjs import RefJS
from oracle import oracle as orc
name = sys.argv[1]; W = int(sys.argv[2]); H = int(sys.argv[3]); P = int(sys.argv[4]) if len(sys.argv) > 4 else 1
t = time.time()
r = RefJS()
info = r.load_test(name)
print(info, 'load %.1fs' % (time.time() - t))
t = time.time()
js = r.scene_json(W, H)
print('json', len(js), '%.1fs' % (time.time() - t))
t = time.time()
mean, rgba, draws = r.render(W, H, P)
print('js render %.1fs' % (time.time() - t), 'draws/sample min/max', draws.min(), draws.max())
sc = orc.OracleScene(js)
simple = info['renderer'] == 'SimpleRenderer'
L = orc.lib()
acc = np.zeros((H, W, 3), dtype=np.float32)
cnt = np.zeros(22, dtype=np.uint64)
flags = 2 | (1 if simple else 0)
npass = 1 if simple else P
rc = L.orc_render(sc._h, W, H, 0, npass, 1, flags, 0, 1, acc.ctypes.data, 1, cnt.ctypes.data)
assert rc == 0, L.orc_last_error()
omean = (acc.astype(np.float64) * (1.0 / npass)).astype(np.float32)
d = np.abs(omean.astype(np.float64) - mean.astype(np.float64))
print('max abs diff', d.max(), 'pixels differing', int((d.max(-1) > 0).sum()), 'of', W * H, ' >1e-6:', int((d.max(-1) > 1e-6).sum()), ' >1e-3:', int((d.max(-1) > 1e-3).sum()))
o8 = orc.resolve_rgba8(acc, npass)
print('rgba8 equal:', np.array_equal(o8, rgba), 'max', np.abs(o8.astype(int) - rgba.astype(int)).max())
if d.max() > 1e-6:
    ys, xs = np.nonzero(d.max(-1) > 1e-6)
    for y, x in list(zip(ys, xs))[:5]:
        print((x, y), mean[y, x], omean[y, x], draws[y, x])
