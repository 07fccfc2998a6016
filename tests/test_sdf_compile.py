"""Host-side checks of the SDF bytecode compiler (csrc/sdf_compile.cpp) that need no GPU: program lengths after the peephole
fusions.  The arithmetic of the fused instructions is checked on the GPU against the oracle, which evaluates the operator
tree node by node like the reference (tests/test_gpu_parity.py: SDF_Menger, SDF_CrossFolds, ...)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _instrs(name, env=None):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from jsraytracer_b200 import lib, scenes\n"
            "from jsraytracer_b200.serializer import Serializer\n"
            "sc = lib.Scene(Serializer(scenes.configure(%r, width=32, height=32)).to_msgpack(), lib.FORMAT_MSGPACK, device=None)\n"
            "print(sc.info['n_sdf_instrs'], sc.info['n_sdfs'])\n") % (ROOT, name)
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, check=True).stdout.split()
    return int(out[0]), int(out[1])


def test_menger_cross_is_one_leaf():
    """DifferenceSDF(Box, RecursiveTransformUnion(Union of three axis bars, [scale, repeat], 6)) (tests/SDF_Menger/test.mjs:27-40):
    BOX; PUSHP; CROSS; 6 x {XFORM; REP; CROSS; MULS_MIN}; POPP; NEG; MAX; END = 31 — 45 before the bars were fused, 77
    without any peephole (JSRT_SDF_FUSE=0 keeps the plain code for A/B runs) — and the six identical steps of the recursion
    are one RTU_CROSS instruction: BOX; PUSHP; CROSS; RTU_CROSS; POPP; NEG; MAX; END = 8."""
    assert _instrs("SDF_Menger") == (8, 1)
    assert _instrs("SDF_Menger", {"JSRT_SDF_RTU": "0"}) == (31, 1)
    assert _instrs("SDF_Menger", {"JSRT_SDF_FUSE": "0"})[0] > 45


def test_cross_folds_and_unequal_bars():
    """CROSS under an intersection (folds with max): BOX; CROSS; END = 3.  CROSS after another child of the same union (folds with
    min), then the intersection's MAX: BOX; SPHERE; CROSS; MAX; END = 5 (+ its material program: a SphereSDF leaf supplies UVs).
    Bars of unequal thickness stay three BoxSDF leaves: BOX; BOX; BOX; BOX; MAX; END = 6."""
    n, progs = _instrs("SDF_CrossFolds")
    plain, _ = _instrs("SDF_CrossFolds", {"JSRT_SDF_FUSE": "0"})
    assert progs == 3
    # the unfused program spends 2 more instructions per fused cross (two BOX) and one per folded leaf (MIN / MAX kept apart)
    assert plain > n
    mat = n - (3 + 5 + 6)
    assert mat > 0 and mat < 40, n
