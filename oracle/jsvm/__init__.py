"""TEST INFRASTRUCTURE — `oracle.jsvm`: a small JavaScript interpreter written for one purpose: executing the
UNMODIFIED sources of the reference (`/root/reference/src/*.js` and its `tests/*/test.mjs` scene files) in a
container without a JS engine, so that the C++ restatement oracle can be pinned against outputs of the
reference itself (tests/golden/refjs_*.npz; generator: oracle/refjs_golden.py).  Never imported by the product.
"""
from .interp import VM                                  # noqa: F401
from .runtime import UNDEF, JSThrow, JSObject, JSArray, JSTypedArray   # noqa: F401
