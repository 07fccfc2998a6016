"""Malformed scene blobs must be rejected with an error, never crash the host process (the C ABI runs inside Node in the
reference's deployment).  Each case found by tools/fuzz_wire.cpp is pinned here; every call runs in a subprocess so that
a regression shows up as a failed test instead of a dead test runner."""
import json
import os
import struct
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


REJECT = """
def rejected(blob, fmt):
    try:
        lib.Scene(blob, fmt, device=None)
    except lib.JsrtError as e:
        return str(e)
    raise SystemExit("accepted a malformed blob")
"""


def _run(body):
    """Runs `body` (python source with `lib`, `scenes`, `Serializer`, `json` in scope) in a fresh interpreter."""
    src = "import sys, json, struct\nsys.path.insert(0, %r)\nfrom jsraytracer_b200 import lib, scenes\nfrom jsraytracer_b200.serializer import Serializer\n" % ROOT
    r = subprocess.run([sys.executable, "-c", src + REJECT + textwrap.dedent(body)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, "exit %d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


def test_cyclic_kdtree_reference_is_rejected():
    out = _run("""
    doc = json.loads(Serializer(scenes.configure("AHollowTetrahedron", width=8, height=8)).to_json())
    def find(o):                                   # first BVHAggregateNode with children
        if isinstance(o, dict):
            v = o.get("_v")
            if isinstance(v, dict) and "lesser_node" in v and v.get("lesser_node"):
                return o
            for x in o.values():
                r = find(x)
                if r: return r
        elif isinstance(o, list):
            for x in o:
                r = find(x)
                if r: return r
    node = find(doc)
    node["_r"] = 987654
    node["_v"]["lesser_node"] = {"_r": 987654}      # the node is its own child
    print(rejected(json.dumps(doc), 0))
    """)
    assert "nested too deep" in out


def test_cyclic_sdf_and_runaway_iterations_are_rejected():
    out = _run("""
    doc = json.loads(Serializer(scenes.configure("SDF_Menger", width=8, height=8)).to_json())
    def find(o, key):
        if isinstance(o, dict):
            v = o.get("_v")
            if isinstance(v, dict) and key in v:
                return o
            for x in o.values():
                r = find(x, key)
                if r: return r
        elif isinstance(o, list):
            for x in o:
                r = find(x, key)
                if r: return r
    rec = find(doc, "iterations")                  # RecursiveTransformUnionSDF
    keep = rec["_v"]["iterations"]
    rec["_v"]["iterations"] = 1e12
    print(rejected(json.dumps(doc), 0))
    rec["_v"]["iterations"] = keep
    rec["_r"] = 424242
    rec["_v"]["sdf"] = {"_r": 424242}              # the union contains itself
    print(rejected(json.dumps(doc), 0))
    """)
    assert "iteration count out of range" in out and "nested too deep" in out


def test_msgpack_count_beyond_the_input_is_rejected():
    out = _run("""
    print(rejected(b"\\xdd\\xff\\xff\\xff\\xff", 1))            # array32 of 4 G elements, no payload
    print(rejected(b"\\xdf\\x7f\\xff\\xff\\xff\\x01", 1))        # map32
    print(rejected(b"\\xc6\\xff\\xff\\xff\\xff\\x00", 1))        # bin32 longer than the blob
    print(rejected(b"[" * 100000, 0))                          # JSON nesting
    print(rejected(b"\\x91" * 100000, 1))                       # msgpack nesting
    print(rejected(b"{}", 0))
    """)
    assert out.count("jsrt:") >= 6


@pytest.mark.parametrize("name", ["Aggregates", "SDF_Combinations", "AHollowTetrahedron", "textured"])
def test_random_mutations_never_crash(name):
    """400 random byte-level mutations per wire format (flip / delete / insert / truncate / splice): every one either
    flattens or is rejected with a message."""
    out = _run("""
    import random
    ser = Serializer(scenes.configure(%r, width=12, height=8))
    for fmt, base in ((1, ser.to_msgpack()), (0, ser.to_json().encode())):
        rng = random.Random(20261018 + fmt)
        ok = bad = 0
        for it in range(400):
            b = bytearray(base)
            mode = rng.random()
            for _ in range(rng.choice([1, 1, 2, 4, 16])):
                if not b: break
                pos = rng.randrange(len(b))
                if mode < 0.45: b[pos] = rng.randrange(256)
                elif mode < 0.6: del b[pos:pos + rng.randrange(1, 64)]
                elif mode < 0.75: b[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
                elif mode < 0.85: del b[pos:]
                else:
                    src = rng.randrange(len(b)); n = rng.randrange(1, 256)
                    chunk = b[src:src + n]; b[pos:pos + len(chunk)] = chunk
            try:
                s = lib.Scene(bytes(b) or b"\\0", fmt, device=None); s.bvh_world_boxes(); s.close(); ok += 1
            except lib.JsrtError:
                bad += 1
        print("fmt", fmt, "accepted", ok, "rejected", bad)
    """ % name)
    assert out.count("accepted") == 2


def test_type_index_and_reference_ids_are_range_checked():
    """`_t` = ["X", 1e9] made the type table resize to ~32 GB; negative / NaN / non-integer indices and
    reference ids were cast unchecked (undefined behaviour)."""
    out = _run("""
    good = json.loads(Serializer(scenes.configure("BoxBall", width=8, height=8)).to_json())
    def first_typed(o):
        if isinstance(o, dict):
            if isinstance(o.get("_t"), list): return o
            for x in o.values():
                r = first_typed(x)
                if r: return r
        elif isinstance(o, list):
            for x in o:
                r = first_typed(x)
                if r: return r
    for bad in (1e9, -1, 1.5, 70000):
        doc = json.loads(json.dumps(good))
        first_typed(doc)["_t"][1] = bad
        print(rejected(json.dumps(doc), 0))
    doc = json.loads(json.dumps(good))
    first_typed(doc)["_t"][0] = 7                        # class name must be a string
    print(rejected(json.dumps(doc), 0))
    doc = json.loads(json.dumps(good))
    doc["_v"]["renderer"] = {"_r": 1e300}                # reference id far outside any integer
    print(rejected(json.dumps(doc), 0))
    """)
    assert out.count("malformed _t") == 5 and "malformed reference _r" in out


def test_json_number_at_the_very_end_of_an_unterminated_buffer():
    """strtod ran on a (pointer, length) blob that need not be NUL-terminated.  A blob whose last token is a
    number, placed at the very end of a page-sized buffer followed by digits that are NOT part of the blob, must parse the
    token as delimited by the length (here: fail cleanly, since the document is truncated) and never read the neighbours."""
    out = _run("""
    import ctypes as C
    L = lib.load()
    text = b'{"width": 12'                                # truncated document ending in a numeric token
    buf = C.create_string_buffer(text + b"3456789" * 8)   # bytes after the blob that look like more digits
    h = L.jsrt_scene_create_host(buf, len(text), 0)
    assert not h
    print(lib.last_error())
    text2 = b'[1.5e3'
    buf2 = C.create_string_buffer(text2 + b"99")
    assert not L.jsrt_scene_create_host(buf2, len(text2), 0)
    print(lib.last_error())
    """)
    assert "jsrt:" in out
