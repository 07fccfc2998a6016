"""TEST INFRASTRUCTURE — value model and type conversions of `oracle/jsvm` (see parser.py for what this package is).

JS value            Python representation
-----------------   -------------------------------------------------------------
undefined           UNDEF (singleton)
null                None
boolean             bool
number              float, always (an IEEE-754 binary64, as in JS)
string              str
object              JSObject (props dict + proto link)
array               JSArray (Python list `items`)
typed array         JSTypedArray (`array.array` of the element type: stores round exactly as JS does)
function / class    JSFunction (interpreted) / NativeFunction (host)
"""
import array as _array
import math
from decimal import Decimal, ROUND_HALF_UP


class Undefined:
    __slots__ = ()

    def __repr__(self):
        return 'undefined'

    def __bool__(self):
        return False


UNDEF = Undefined()


class JSThrow(Exception):
    """a JS `throw`: carries the thrown value"""

    def __init__(self, value):
        Exception.__init__(self)
        self.value = value

    def __str__(self):
        v = self.value
        if isinstance(v, JSObject):
            m = v.props.get('message')
            s = v.props.get('stack')
            return 'JS exception: %s %s' % (m, s or '')
        return 'JS exception: %r' % (v,)


class JSObject:
    __slots__ = ('props', 'proto', 'hidden', 'cls')

    def __init__(self, proto=None):
        self.props = {}
        self.proto = proto
        self.hidden = None     # set of non-enumerable own keys (Object.defineProperty enumerable:false)
        self.cls = 'Object'


class JSArray(JSObject):
    __slots__ = ('items',)

    def __init__(self, proto, items):
        self.props = {}
        self.proto = proto
        self.hidden = None
        self.cls = 'Array'
        self.items = items


TYPED_KINDS = {
    # name: (array typecode, bytes, integer?, lo, hi, clamped)
    'Float32Array': ('f', 4, False, 0, 0, False),
    'Float64Array': ('d', 8, False, 0, 0, False),
    'Int8Array': ('b', 1, True, -128, 127, False),
    'Uint8Array': ('B', 1, True, 0, 255, False),
    'Uint8ClampedArray': ('B', 1, True, 0, 255, True),
    'Int16Array': ('h', 2, True, -32768, 32767, False),
    'Uint16Array': ('H', 2, True, 0, 65535, False),
    'Int32Array': ('i', 4, True, -2 ** 31, 2 ** 31 - 1, False),
    'Uint32Array': ('I', 4, True, 0, 2 ** 32 - 1, False),
}


class JSTypedArray(JSObject):
    __slots__ = ('items', 'kind', 'isint', 'lo', 'span', 'clamped')

    def __init__(self, proto, kind, n_or_items):
        self.props = {}
        self.proto = proto
        self.hidden = None
        self.cls = kind
        self.kind = kind
        code, _, isint, lo, hi, clamped = TYPED_KINDS[kind]
        self.isint = isint
        self.lo = lo
        self.span = hi - lo + 1
        self.clamped = clamped
        if isinstance(n_or_items, int):
            self.items = _array.array(code, bytes(n_or_items * _array.array(code).itemsize))
        else:
            self.items = _array.array(code)
            for v in n_or_items:
                self.items.append(self.conv(v))

    def conv(self, v):
        """JS number -> stored element (floats are rounded by array.array itself)"""
        if not self.isint:
            return v
        if v != v or v in (math.inf, -math.inf):
            return 0
        if self.clamped:
            if v <= 0:
                return 0
            if v >= 255:
                return 255
            r = math.floor(v)
            d = v - r
            if d > 0.5 or (d == 0.5 and r % 2 == 1):
                r += 1
            return int(r)
        i = int(v)   # truncation toward zero
        return (i - self.lo) % self.span + self.lo


class JSArrayBuffer(JSObject):
    """`typedArray.buffer` / `new ArrayBuffer(n)`: the bytes of one owning typed array.  A view constructed over it
    (`new Uint8Array(buffer, offset, length)`) is a COPY of those bytes — enough for reading, which is all the code run
    here does with such views."""
    __slots__ = ('owner',)


class JSFunction(JSObject):
    __slots__ = ('name', 'env', 'pnames', 'binder', 'body', 'is_arrow', 'is_gen', 'expr_body', 'home', 'is_class',
                 'parent', 'ctor', 'fields', 'uses_args', 'nparams', 'vm', 'var_names', 'is_derived')


class NativeFunction(JSObject):
    __slots__ = ('name', 'fn', 'ctor')

    def __init__(self, proto, name, fn, ctor=None):
        self.props = {}
        self.proto = proto
        self.hidden = None
        self.cls = 'Function'
        self.name = name
        self.fn = fn          # fn(this, args) -> value
        self.ctor = ctor      # ctor(args, new_target) -> object, or None when not constructible


class BoundFunction(JSObject):
    __slots__ = ('target', 'this', 'args')


class Scope:
    __slots__ = ('vars', 'parent')

    def __init__(self, parent):
        self.vars = {}
        self.parent = parent


# -- conversions ----------------------------------------------------------------------------------

def typeof(v):
    c = v.__class__
    if c is float:
        return 'number'
    if c is str:
        return 'string'
    if c is bool:
        return 'boolean'
    if v is UNDEF:
        return 'undefined'
    if v is None:
        return 'object'
    if c is JSFunction or c is NativeFunction or c is BoundFunction:
        return 'function'
    return 'object'


def truthy(v):
    c = v.__class__
    if c is bool:
        return v
    if c is float:
        return v == v and v != 0.0
    if c is str:
        return v != ''
    return v is not None and v is not UNDEF


def num_to_str(v):
    """Number::toString (radix 10)"""
    if v != v:
        return 'NaN'
    if v == math.inf:
        return 'Infinity'
    if v == -math.inf:
        return '-Infinity'
    if v == 0:
        return '0'
    if abs(v) < 9007199254740992.0 and v == int(v):
        return str(int(v))
    r = repr(abs(v))
    sign = '-' if v < 0 else ''
    # shortest round-trip digits and decimal exponent
    if 'e' in r:
        mant, ex = r.split('e')
        ex = int(ex)
    else:
        mant, ex = r, 0
    if '.' in mant:
        ip, fp = mant.split('.')
    else:
        ip, fp = mant, ''
    if fp == '0':
        fp = ''
    digits = (ip + fp).lstrip('0')
    # n: position of the decimal point relative to the first digit
    n = len(ip.lstrip('0')) + ex if ip.strip('0') else ex - (len(fp) - len(fp.lstrip('0')))
    digits = digits.rstrip('0') or '0'
    k = len(digits)
    if k <= n <= 21:
        return sign + digits + '0' * (n - k)
    if 0 < n <= 21:
        return sign + digits[:n] + '.' + digits[n:]
    if -6 < n <= 0:
        return sign + '0.' + '0' * (-n) + digits
    e = n - 1
    es = ('+' if e >= 0 else '-') + str(abs(e))
    if k == 1:
        return sign + digits + 'e' + es
    return sign + digits[0] + '.' + digits[1:] + 'e' + es


def str_to_num(s):
    s = s.strip()
    if s == '':
        return 0.0
    try:
        if s[:2] in ('0x', '0X'):
            return float(int(s, 16))
        if s in ('Infinity', '+Infinity'):
            return math.inf
        if s == '-Infinity':
            return -math.inf
        if s.lower().lstrip('+-') in ('inf', 'infinity', 'nan') or '_' in s:
            return math.nan
        return float(s)
    except ValueError:
        return math.nan


def to_precision(v, p):
    """Number.prototype.toPrecision: exact decimal expansion of the double, ties away from zero (the spec's 'larger n')"""
    if v != v:
        return 'NaN'
    if v in (math.inf, -math.inf):
        return 'Infinity' if v > 0 else '-Infinity'
    p = int(p)
    if v == 0:
        return '0' if p == 1 else '0.' + '0' * (p - 1)
    sign = '-' if v < 0 else ''
    d = Decimal(abs(v))
    e = d.adjusted()
    q = Decimal(1).scaleb(e - p + 1)
    r = d.quantize(q, rounding=ROUND_HALF_UP)
    if r.adjusted() != e:          # rounding carried into a new digit
        e = r.adjusted()
        q = Decimal(1).scaleb(e - p + 1)
        r = d.quantize(q, rounding=ROUND_HALF_UP)
    digits = str(int(r.scaleb(-(e - p + 1))))
    if e < -6 or e >= p:
        m = digits[0] + ('.' + digits[1:] if p > 1 else '')
        return sign + m + 'e' + ('+' if e >= 0 else '-') + str(abs(e))
    if e >= 0:
        return sign + digits[:e + 1] + ('.' + digits[e + 1:] if p > e + 1 else '')
    return sign + '0.' + '0' * (-e - 1) + digits


def to_fixed(v, n):
    if v != v:
        return 'NaN'
    if abs(v) >= 1e21:
        return num_to_str(v)
    n = int(n)
    d = Decimal(v).quantize(Decimal(1).scaleb(-n), rounding=ROUND_HALF_UP)
    s = format(d, 'f')
    if d == 0 and s.startswith('-') and v >= 0:
        s = s[1:]
    return s


def to_int32(v):
    if v != v or v in (math.inf, -math.inf):
        return 0
    i = int(v) & 0xFFFFFFFF
    return i - 0x100000000 if i & 0x80000000 else i


def to_uint32(v):
    if v != v or v in (math.inf, -math.inf):
        return 0
    return int(v) & 0xFFFFFFFF
