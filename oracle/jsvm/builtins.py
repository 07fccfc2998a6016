"""TEST INFRASTRUCTURE — the standard built-in objects of `oracle/jsvm` (the part of ECMAScript the reference uses).

Numeric functions are the host libm's (Python `math`), the same library the C++ oracle links; V8's own
fdlibm port may differ from it in the last bit of sin / cos / pow for some arguments.
"""
import math
import re
import sys
import time

from .runtime import (UNDEF, JSThrow, JSObject, JSArray, JSTypedArray, JSArrayBuffer, JSFunction, NativeFunction, BoundFunction,
                      TYPED_KINDS, typeof, truthy, num_to_str, str_to_num, to_precision, to_fixed)


class JSRegExp(JSObject):
    __slots__ = ('rx', 'source', 'flags', 'last_index')


class JSMap(JSObject):
    __slots__ = ('data',)



class JSSet(JSObject):
    __slots__ = ('data',)


class JSPromise(JSObject):
    __slots__ = ('state', 'value', 'reactions')


class JSListIter(JSObject):
    """iterator object over a host list (Array.prototype.entries / keys / values, Map iterators)"""
    __slots__ = ('it',)

    def py_iter(self):
        return self.it


def _key(v):
    """Map / Set key normalisation (SameValueZero): floats by value, everything else by identity or equality of str"""
    if v.__class__ is float:
        if v != v:
            return ('nan',)
        return v + 0.0
    if v.__class__ is bool:
        return ('bool', v)
    return v


def js_regex_to_py(src, flags):
    f = 0
    if 'i' in flags:
        f |= re.IGNORECASE
    if 'm' in flags:
        f |= re.MULTILINE
    if 's' in flags:
        f |= re.DOTALL
    # JS-only syntax used rarely: named groups (?<n>...) -> (?P<n>...)
    src = re.sub(r'\(\?<([A-Za-z_]\w*)>', r'(?P<\1>', src)
    # `$` in JS without the m flag matches only at the very end
    return re.compile(src, f)


def install(vm):
    G = vm.root.vars
    ObjectProto = JSObject(None)
    FunctionProto = JSObject(ObjectProto)
    vm.ObjectProto = ObjectProto
    vm.FunctionProto = FunctionProto

    def native(name, fn, ctor=None):
        return NativeFunction(FunctionProto, name, fn, ctor)

    def method(obj, name, fn):
        obj.props[name] = native(name, fn)
        if obj.hidden is None:
            obj.hidden = set()
        obj.hidden.add(name)

    def arg(args, i):
        return args[i] if i < len(args) else UNDEF

    def make_ctor(name, proto, call_fn, ctor_fn, parent_ctor=None):
        c = native(name, call_fn, ctor_fn)
        if parent_ctor is not None:
            c.proto = parent_ctor
        c.props['prototype'] = proto
        c.hidden = {'prototype'}
        proto.props['constructor'] = c
        if proto.hidden is None:
            proto.hidden = set()
        proto.hidden.add('constructor')
        G[name] = c
        return c

    def proto_of(nt, default):
        p = vm.get(nt, 'prototype')
        return p if isinstance(p, JSObject) else default

    tostr, tonum, get, put, call = vm.tostr, vm.tonum, vm.get, vm.put, vm.call

    def is_callable(f):
        return isinstance(f, (JSFunction, NativeFunction, BoundFunction))

    def need_fn(f):
        if not is_callable(f):
            vm.throw('TypeError', '%s is not a function' % (tostr(f) if not isinstance(f, JSObject) else 'object'))
        return f

    def to_index(v, n, default):
        """relative index argument (slice / splice / fill ...)"""
        if v is UNDEF:
            return default
        x = tonum(v)
        if x != x:
            return 0
        if x < 0:
            return max(0, n + int(max(x, -1e18)))
        return int(min(x, n))

    # ------------------------------------------------------------------------------------------
    # global object
    gobj = JSObject(ObjectProto)
    gobj.props = G          # the global scope and the global object share their bindings
    vm.global_object = gobj
    for name in ('globalThis', 'window', 'self'):
        G[name] = gobj
    G['undefined'] = UNDEF
    G['NaN'] = math.nan
    G['Infinity'] = math.inf

    # ------------------------------------------------------------------------------------------
    # Object
    def object_call(this, args):
        v = arg(args, 0)
        if v is None or v is UNDEF:
            return JSObject(ObjectProto)
        return v

    Object = make_ctor('Object', ObjectProto, object_call, lambda args, nt: object_call(None, args) if nt is Object else JSObject(proto_of(nt, ObjectProto)))

    def object_assign(this, args):
        target = arg(args, 0)
        for src in args[1:]:
            if isinstance(src, (JSObject, str)):
                for k in vm.own_keys(src):
                    put(target, k, get(src, k))
        return target

    def object_define_property(this, args):
        o, k, desc = arg(args, 0), vm.tokey(arg(args, 1)), arg(args, 2)
        if get(desc, 'get') is not UNDEF or get(desc, 'set') is not UNDEF:
            raise NotImplementedError('accessor properties')
        o.props[k] = get(desc, 'value')
        if not truthy(get(desc, 'enumerable')):
            if o.hidden is None:
                o.hidden = set()
            o.hidden.add(k)
        elif o.hidden is not None:
            o.hidden.discard(k)
        return o

    def arr(items):
        return JSArray(vm.ArrayProto, items)

    Object.props.update({
        'assign': native('assign', object_assign),
        'defineProperty': native('defineProperty', object_define_property),
        'keys': native('keys', lambda this, a: arr(list(vm.own_keys(arg(a, 0))))),
        'values': native('values', lambda this, a: arr([get(arg(a, 0), k) for k in vm.own_keys(arg(a, 0))])),
        'entries': native('entries', lambda this, a: arr([arr([k, get(arg(a, 0), k)]) for k in vm.own_keys(arg(a, 0))])),
        'getPrototypeOf': native('getPrototypeOf', lambda this, a: arg(a, 0).proto if isinstance(arg(a, 0), JSObject) else None),
        'setPrototypeOf': native('setPrototypeOf', lambda this, a: (setattr(arg(a, 0), 'proto', arg(a, 1)), arg(a, 0))[1]),
        'create': native('create', lambda this, a: JSObject(arg(a, 0) if isinstance(arg(a, 0), JSObject) else None)),
        'freeze': native('freeze', lambda this, a: arg(a, 0)),
        'fromEntries': native('fromEntries', lambda this, a: _from_entries(a)),
        'getOwnPropertyNames': native('getOwnPropertyNames', lambda this, a: arr(list(arg(a, 0).props.keys()))),
    })

    def _from_entries(a):
        o = JSObject(ObjectProto)
        for e in vm.iterate(arg(a, 0)):
            o.props[vm.tokey(get(e, 0.0))] = get(e, 1.0)
        return o

    def object_to_string(this, args):
        if isinstance(this, JSObject):
            return '[object %s]' % ('Array' if this.__class__ is JSArray else 'Function' if is_callable(this) else 'Object')
        return '[object %s]' % ('Undefined' if this is UNDEF else 'Null' if this is None else typeof(this).capitalize())

    method(ObjectProto, 'toString', object_to_string)
    method(ObjectProto, 'valueOf', lambda this, a: this)
    method(ObjectProto, 'hasOwnProperty', lambda this, a: _has_own(this, arg(a, 0)))
    method(ObjectProto, 'isPrototypeOf', lambda this, a: _is_proto_of(this, arg(a, 0)))

    def _has_own(o, k):
        c = o.__class__
        if c is JSArray or c is JSTypedArray:
            if k.__class__ is float:
                return k == int(k) and 0 <= k < len(o.items)
            if k == 'length':
                return True
        return isinstance(o, JSObject) and vm.tokey(k) in o.props

    def _is_proto_of(p, o):
        o = o.proto if isinstance(o, JSObject) else None
        while o is not None:
            if o is p:
                return True
            o = o.proto
        return False

    # ------------------------------------------------------------------------------------------
    # Function
    def function_ctor(args, nt):
        params = ','.join(tostr(a) for a in args[:-1])
        body = tostr(args[-1]) if args else ''
        return vm.make_function_from_source(params, body)

    Function = make_ctor('Function', FunctionProto, lambda this, a: function_ctor(a, None), function_ctor)

    def fn_apply(this, args):
        a = arg(args, 1)
        return call(need_fn(this), arg(args, 0), [] if (a is None or a is UNDEF) else list(vm.iterate(a)))

    def fn_bind(this, args):
        b = BoundFunction.__new__(BoundFunction)
        b.props = {}
        b.proto = FunctionProto
        b.hidden = None
        b.cls = 'Function'
        b.target = need_fn(this)
        b.this = arg(args, 0)
        b.args = list(args[1:])
        return b

    method(FunctionProto, 'call', lambda this, a: call(need_fn(this), arg(a, 0), list(a[1:])))
    method(FunctionProto, 'apply', fn_apply)
    method(FunctionProto, 'bind', fn_bind)
    method(FunctionProto, 'toString', lambda this, a: 'function %s() { [code] }' % getattr(this, 'name', ''))

    # ------------------------------------------------------------------------------------------
    # Error types
    def make_error_type(name, parent_proto, parent_ctor):
        proto = JSObject(parent_proto)
        proto.props['name'] = name
        proto.props['message'] = ''

        def ctor(args, nt):
            e = JSObject(proto_of(nt, proto))
            if arg(args, 0) is not UNDEF:
                e.props['message'] = tostr(arg(args, 0))
            e.props['stack'] = ''
            e.hidden = {'stack', 'message'}
            return e
        c = make_ctor(name, proto, lambda this, a: ctor(a, c), ctor, parent_ctor)
        return c, proto

    Error, ErrorProto = make_error_type('Error', ObjectProto, None)
    method(ErrorProto, 'toString', lambda this, a: '%s: %s' % (tostr(get(this, 'name')), tostr(get(this, 'message'))))
    for n in ('TypeError', 'RangeError', 'ReferenceError', 'SyntaxError', 'EvalError'):
        make_error_type(n, ErrorProto, Error)

    # ------------------------------------------------------------------------------------------
    # Array
    ArrayProto = JSArray(ObjectProto, [])
    vm.ArrayProto = ArrayProto

    def array_ctor(args, nt):
        proto = proto_of(nt, ArrayProto)
        if len(args) == 1 and args[0].__class__ is float:
            n = args[0]
            if n != int(n) or n < 0 or n > 4294967295:
                vm.throw('RangeError', 'invalid array length')
            return JSArray(proto, [UNDEF] * int(n))
        return JSArray(proto, list(args))

    Array = make_ctor('Array', ArrayProto, lambda this, a: array_ctor(a, Array), array_ctor)

    def species_new(exemplar, n):
        """ArraySpeciesCreate: subclasses of Array get `new C(n)` (this is how `Mat` behaves under map / slice)"""
        C = get(exemplar, 'constructor')
        if C is Array or C is UNDEF or not isinstance(C, JSObject):
            return JSArray(ArrayProto, [])
        return vm.construct(C, [float(n)])

    def ctor_new(C, args):
        if C is Array or not is_callable(C):
            return array_ctor(args, Array)
        return vm.construct(C, args)

    def set_at(a, i, v):
        """CreateDataProperty on an array-like under construction"""
        if a.__class__ is JSArray:
            items = a.items
            if i < len(items):
                items[i] = v
            else:
                items.extend([UNDEF] * (i - len(items)))
                items.append(v)
        else:
            put(a, float(i), v)

    def set_len(a, n):
        if a.__class__ is JSArray:
            items = a.items
            if n < len(items):
                del items[n:]
            else:
                items.extend([UNDEF] * (n - len(items)))
        else:
            put(a, 'length', float(n))

    def items_of(o):
        """list view of an array-like `this`"""
        c = o.__class__
        if c is JSArray:
            return o.items
        if c is JSTypedArray:
            return [float(x) for x in o.items] if o.isint else o.items
        if c is str:
            return list(o)
        n = int(tonum(get(o, 'length')))
        return [get(o, float(i)) for i in range(n)]

    def array_from(this, args):
        src, fn = arg(args, 0), arg(args, 1)
        c = src.__class__
        if c in (JSArray, JSTypedArray, str) or hasattr(src, 'py_iter') or hasattr(src, 'pygen'):
            vals = list(vm.iterate(src))
            out = ctor_new(this, [])
        else:
            vals = items_of(src) if isinstance(src, JSObject) else []
            out = ctor_new(this, [float(len(vals))])
        if fn is not UNDEF:
            vals = [call(fn, UNDEF, [v, float(i)]) for i, v in enumerate(vals)]
        for i, v in enumerate(vals):
            set_at(out, i, v)
        set_len(out, len(vals))
        return out

    def array_of(this, args):
        out = ctor_new(this, [float(len(args))])
        for i, v in enumerate(args):
            set_at(out, i, v)
        set_len(out, len(args))
        return out

    Array.props['isArray'] = native('isArray', lambda this, a: arg(a, 0).__class__ is JSArray)
    Array.props['from'] = native('from', array_from)
    Array.props['of'] = native('of', array_of)

    def a_push(this, args):
        if this.__class__ is JSArray:
            this.items.extend(args)
            return float(len(this.items))
        n = int(tonum(get(this, 'length')))
        for v in args:
            put(this, float(n), v)
            n += 1
        put(this, 'length', float(n))
        return float(n)

    def a_pop(this, args):
        return this.items.pop() if this.items else UNDEF

    def a_shift(this, args):
        return this.items.pop(0) if this.items else UNDEF

    def a_unshift(this, args):
        this.items[0:0] = args
        return float(len(this.items))

    def a_splice(this, args):
        items = this.items
        n = len(items)
        start = to_index(arg(args, 0), n, 0)
        if len(args) == 0:
            cnt = 0
        elif len(args) == 1:
            cnt = n - start
        else:
            c = tonum(args[1])
            cnt = int(min(max(0 if c != c else c, 0), n - start))
        removed = items[start:start + cnt]
        out = species_new(this, cnt)
        for i, v in enumerate(removed):
            set_at(out, i, v)
        set_len(out, cnt)
        items[start:start + cnt] = list(args[2:])
        return out

    def a_slice(this, args):
        items = items_of(this)
        n = len(items)
        s = to_index(arg(args, 0), n, 0)
        e = to_index(arg(args, 1), n, n)
        part = items[s:e] if e > s else []
        out = species_new(this, len(part))
        for i, v in enumerate(part):
            set_at(out, i, v)
        set_len(out, len(part))
        return out

    def a_concat(this, args):
        out = species_new(this, 0)
        k = 0
        for x in [this] + list(args):
            if x.__class__ is JSArray:
                for v in x.items:
                    set_at(out, k, v)
                    k += 1
            else:
                set_at(out, k, x)
                k += 1
        set_len(out, k)
        return out

    def a_map(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        items = items_of(this)
        n = len(items)
        out = species_new(this, n)
        i = 0
        if fn.__class__ is JSFunction and fn.pnames is not None and len(fn.pnames) == 1 and not fn.uses_args:
            while i < n and i < len(items):
                set_at(out, i, call(fn, that, [items[i]]))
                i += 1
        else:
            while i < n and i < len(items):
                set_at(out, i, call(fn, that, [items[i], float(i), this]))
                i += 1
        return out

    def a_for_each(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        items = items_of(this)
        n = len(items)
        i = 0
        while i < n and i < len(items):
            call(fn, that, [items[i], float(i), this])
            i += 1
        return UNDEF

    def a_filter(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        items = items_of(this)
        out = species_new(this, 0)
        k = 0
        for i, v in enumerate(list(items)):
            if truthy(call(fn, that, [v, float(i), this])):
                set_at(out, k, v)
                k += 1
        return out

    def a_every(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        items = items_of(this)
        i = 0
        while i < len(items):
            if not truthy(call(fn, that, [items[i], float(i), this])):
                return False
            i += 1
        return True

    def a_some(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        items = items_of(this)
        i = 0
        while i < len(items):
            if truthy(call(fn, that, [items[i], float(i), this])):
                return True
            i += 1
        return False

    def a_find(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        for i, v in enumerate(list(items_of(this))):
            if truthy(call(fn, that, [v, float(i), this])):
                return v
        return UNDEF

    def a_find_index(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        for i, v in enumerate(list(items_of(this))):
            if truthy(call(fn, that, [v, float(i), this])):
                return float(i)
        return -1.0

    def a_reduce(this, args):
        fn = need_fn(arg(args, 0))
        items = items_of(this)
        i = 0
        if len(args) >= 2:
            acc = args[1]
        else:
            if not len(items):
                vm.throw('TypeError', 'reduce of empty array with no initial value')
            acc = items[0]
            i = 1
        while i < len(items):
            acc = call(fn, UNDEF, [acc, items[i], float(i), this])
            i += 1
        return acc

    def a_reduce_right(this, args):
        fn = need_fn(arg(args, 0))
        items = list(items_of(this))
        i = len(items) - 1
        if len(args) >= 2:
            acc = args[1]
        else:
            if not items:
                vm.throw('TypeError', 'reduce of empty array with no initial value')
            acc = items[i]
            i -= 1
        while i >= 0:
            acc = call(fn, UNDEF, [acc, items[i], float(i), this])
            i -= 1
        return acc

    def seq(a, b):
        ca = a.__class__
        if ca is float:
            return b.__class__ is float and a == b
        if ca is str:
            return b.__class__ is str and a == b
        return a is b

    def a_index_of(this, args):
        x = arg(args, 0)
        items = items_of(this)
        start = to_index(arg(args, 1), len(items), 0)
        for i in range(start, len(items)):
            if seq(items[i], x):
                return float(i)
        return -1.0

    def a_last_index_of(this, args):
        x = arg(args, 0)
        items = items_of(this)
        for i in range(len(items) - 1, -1, -1):
            if seq(items[i], x):
                return float(i)
        return -1.0

    def a_includes(this, args):
        x = arg(args, 0)
        for v in items_of(this):
            if seq(v, x) or (x.__class__ is float and x != x and v.__class__ is float and v != v):
                return True
        return False

    def a_join(this, args):
        sep = ',' if arg(args, 0) is UNDEF else tostr(args[0])
        return sep.join('' if (v is None or v is UNDEF) else tostr(v) for v in items_of(this))

    def a_reverse(this, args):
        this.items.reverse()
        return this

    def a_fill(this, args):
        items = this.items
        n = len(items)
        v = arg(args, 0)
        s = to_index(arg(args, 1), n, 0)
        e = to_index(arg(args, 2), n, n)
        if this.__class__ is JSTypedArray:
            v = this.conv(tonum(v))
        for i in range(s, e):
            items[i] = v
        return this

    def a_sort(this, args):
        import functools
        fn = arg(args, 0)
        items = this.items
        undef = [v for v in items if v is UNDEF]
        rest = [v for v in items if v is not UNDEF]
        if fn is UNDEF:
            if this.__class__ is JSTypedArray:
                rest.sort()
            else:
                rest.sort(key=lambda v: [ord(ch) for ch in tostr(v)])
        else:
            def cmp(a, b):
                r = tonum(call(fn, UNDEF, [float(a) if a.__class__ is int else a, float(b) if b.__class__ is int else b]))
                return -1 if r < 0 else (1 if r > 0 else 0)
            rest.sort(key=functools.cmp_to_key(cmp))    # stable, as Array.prototype.sort has been since ES2019
        for i, v in enumerate(rest + undef):
            items[i] = v
        return this

    def a_flat(this, args):
        depth = 1 if arg(args, 0) is UNDEF else int(tonum(args[0]))
        out = []

        def rec(a, d):
            for v in a.items:
                if v.__class__ is JSArray and d > 0:
                    rec(v, d - 1)
                else:
                    out.append(v)
        rec(this, depth)
        return arr(out)

    def a_flat_map(this, args):
        fn = need_fn(arg(args, 0))
        out = []
        for i, v in enumerate(list(this.items)):
            r = call(fn, arg(args, 1), [v, float(i), this])
            if r.__class__ is JSArray:
                out.extend(r.items)
            else:
                out.append(r)
        return arr(out)

    ListIterProto = JSObject(ObjectProto)

    def list_iter(values):
        it = JSListIter(ListIterProto)
        it.it = iter(values)
        return it

    for name, fn in [('push', a_push), ('pop', a_pop), ('shift', a_shift), ('unshift', a_unshift), ('splice', a_splice),
                     ('slice', a_slice), ('concat', a_concat), ('map', a_map), ('forEach', a_for_each), ('filter', a_filter),
                     ('every', a_every), ('some', a_some), ('find', a_find), ('findIndex', a_find_index),
                     ('reduce', a_reduce), ('reduceRight', a_reduce_right), ('indexOf', a_index_of),
                     ('lastIndexOf', a_last_index_of), ('includes', a_includes), ('join', a_join), ('reverse', a_reverse),
                     ('fill', a_fill), ('sort', a_sort), ('flat', a_flat), ('flatMap', a_flat_map),
                     ('toString', lambda this, a: a_join(this, [])),
                     ('keys', lambda this, a: list_iter([float(i) for i in range(len(this.items))])),
                     ('values', lambda this, a: list_iter(list(items_of(this)))),
                     ('entries', lambda this, a: list_iter([arr([float(i), v]) for i, v in enumerate(items_of(this))]))]:
        method(ArrayProto, name, fn)

    # ------------------------------------------------------------------------------------------
    # typed arrays
    TypedArrayProto = JSObject(ObjectProto)

    def typed_species(exemplar, n):
        C = get(exemplar, 'constructor')
        r = vm.construct(C, [float(n)])
        if r.__class__ is not JSTypedArray:
            vm.throw('TypeError', 'species constructor did not return a typed array')
        return r

    def t_map(this, args):
        fn, that = need_fn(arg(args, 0)), arg(args, 1)
        src = this.items
        n = len(src)
        out = typed_species(this, n)
        dst = out.items
        isint_src = this.isint
        conv = out.conv if out.isint else None
        if fn.__class__ is JSFunction and fn.pnames is not None and not fn.uses_args and not fn.is_gen:
            np_ = len(fn.pnames)
        else:
            np_ = 3
        for i in range(n):
            x = src[i]
            if isint_src:
                x = float(x)
            if np_ == 1:
                r = call(fn, that, [x])
            elif np_ == 2:
                r = call(fn, that, [x, float(i)])
            else:
                r = call(fn, that, [x, float(i), this])
            if r.__class__ is not float:
                r = tonum(r)
            dst[i] = conv(r) if conv is not None else r
        return out

    def t_slice(this, args):
        n = len(this.items)
        s = to_index(arg(args, 0), n, 0)
        e = to_index(arg(args, 1), n, n)
        cnt = max(e - s, 0)
        out = typed_species(this, cnt)
        for i in range(cnt):
            out.items[i] = this.items[s + i]
        return out

    def t_set(this, args):
        src = items_of(arg(args, 0))
        off = 0 if arg(args, 1) is UNDEF else int(tonum(args[1]))
        if off + len(src) > len(this.items):
            vm.throw('RangeError', 'offset is out of bounds')
        for i, v in enumerate(src):
            this.items[off + i] = this.conv(tonum(v)) if this.isint else tonum(v)
        return UNDEF

    def t_filter(this, args):
        fn = need_fn(arg(args, 0))
        kept = [v for i, v in enumerate(items_of(this)) if truthy(call(fn, arg(args, 1), [v, float(i), this]))]
        out = typed_species(this, len(kept))
        for i, v in enumerate(kept):
            out.items[i] = out.conv(v) if out.isint else v
        return out

    for name, fn in [('map', t_map), ('forEach', a_for_each), ('every', a_every), ('some', a_some), ('reduce', a_reduce),
                     ('reduceRight', a_reduce_right), ('fill', a_fill), ('slice', t_slice), ('subarray', t_slice), ('set', t_set),
                     ('join', a_join), ('indexOf', a_index_of), ('lastIndexOf', a_last_index_of), ('includes', a_includes),
                     ('find', a_find), ('findIndex', a_find_index), ('filter', t_filter), ('reverse', a_reverse), ('sort', a_sort),
                     ('toString', lambda this, a: a_join(this, [])),
                     ('keys', lambda this, a: list_iter([float(i) for i in range(len(this.items))])),
                     ('values', lambda this, a: list_iter(list(items_of(this)))),
                     ('entries', lambda this, a: list_iter([arr([float(i), v]) for i, v in enumerate(items_of(this))]))]:
        method(TypedArrayProto, name, fn)

    def make_typed(kind):
        proto = JSObject(TypedArrayProto)

        def ctor(args, nt):
            p = proto_of(nt, proto)
            a0 = arg(args, 0)
            if a0 is UNDEF:
                return JSTypedArray(p, kind, 0)
            if a0.__class__ is float:
                if a0 != int(a0) or a0 < 0:
                    vm.throw('RangeError', 'invalid typed array length')
                return JSTypedArray(p, kind, int(a0))
            if a0.__class__ is JSArrayBuffer:
                raw = a0.owner.items.tobytes()
                size = TYPED_KINDS[kind][1]
                off = 0 if arg(args, 1) is UNDEF else int(tonum(args[1]))
                n = (len(raw) - off) // size if arg(args, 2) is UNDEF else int(tonum(args[2]))
                if off % size or off + n * size > len(raw) or n < 0:
                    vm.throw('RangeError', 'invalid typed array offset / length')
                t = JSTypedArray(p, kind, 0)
                t.items.frombytes(raw[off:off + n * size])
                return t
            if isinstance(a0, JSObject):
                if a0.__class__ in (JSArray, JSTypedArray) or hasattr(a0, 'py_iter') or hasattr(a0, 'pygen'):
                    vals = list(vm.iterate(a0))
                else:
                    vals = items_of(a0)
                return JSTypedArray(p, kind, [v if v.__class__ is float else tonum(v) for v in vals])
            return JSTypedArray(p, kind, int(tonum(a0)))

        def call_fn(this, args):
            vm.throw('TypeError', "constructor %s requires 'new'" % kind)

        C = make_ctor(kind, proto, call_fn, ctor)
        C.props['BYTES_PER_ELEMENT'] = float(TYPED_KINDS[kind][1])

        def t_from(this, args):
            src, fn = arg(args, 0), arg(args, 1)
            if src.__class__ in (JSArray, JSTypedArray, str) or hasattr(src, 'py_iter') or hasattr(src, 'pygen'):
                vals = list(vm.iterate(src))
            else:
                vals = items_of(src)
            if fn is not UNDEF:
                vals = [call(fn, arg(args, 2), [v, float(i)]) for i, v in enumerate(vals)]
            out = vm.construct(this, [float(len(vals))])
            dst = out.items
            for i, v in enumerate(vals):
                if v.__class__ is not float:
                    v = tonum(v)
                dst[i] = out.conv(v) if out.isint else v
            return out

        def t_of(this, args):
            out = vm.construct(this, [float(len(args))])
            dst = out.items
            i = 0
            for v in args:
                if v.__class__ is not float:
                    v = tonum(v)
                dst[i] = out.conv(v) if out.isint else v
                i += 1
            return out

        C.props['from'] = native('from', t_from)
        C.props['of'] = native('of', t_of)

    for kind in TYPED_KINDS:
        make_typed(kind)

    ArrayBufferProto = JSObject(ObjectProto)

    def array_buffer_ctor(args, nt):
        b = JSArrayBuffer(proto_of(nt, ArrayBufferProto))
        b.owner = JSTypedArray(G['Uint8Array'].props['prototype'], 'Uint8Array', int(tonum(arg(args, 0))))
        return b
    make_ctor('ArrayBuffer', ArrayBufferProto, lambda this, a: vm.throw('TypeError', "constructor ArrayBuffer requires 'new'"), array_buffer_ctor)

    # ------------------------------------------------------------------------------------------
    # String / Number / Boolean
    StringProto = JSObject(ObjectProto)
    NumberProto = JSObject(ObjectProto)
    BooleanProto = JSObject(ObjectProto)
    vm.StringProto, vm.NumberProto, vm.BooleanProto = StringProto, NumberProto, BooleanProto

    String = make_ctor('String', StringProto, lambda this, a: tostr(a[0]) if a else '', lambda a, nt: tostr(a[0]) if a else '')
    Number = make_ctor('Number', NumberProto, lambda this, a: tonum(a[0]) if a else 0.0, lambda a, nt: tonum(a[0]) if a else 0.0)
    make_ctor('Boolean', BooleanProto, lambda this, a: truthy(arg(a, 0)), lambda a, nt: truthy(arg(a, 0)))
    String.props['fromCharCode'] = native('fromCharCode', lambda this, a: ''.join(chr(int(tonum(x)) & 0xFFFF) for x in a))

    def parse_float(this, args):
        s = tostr(arg(args, 0)).strip()
        m = re.match(r'[+-]?(?:Infinity|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)', s)
        if not m:
            return math.nan
        return str_to_num(m.group(0))

    def parse_int(this, args):
        s = tostr(arg(args, 0)).strip()
        radix = 0 if arg(args, 1) is UNDEF else int(tonum(args[1]))
        sign = 1
        if s[:1] in ('+', '-'):
            if s[0] == '-':
                sign = -1
            s = s[1:]
        if radix in (0, 16) and s[:2] in ('0x', '0X'):
            s = s[2:]
            radix = 16
        if radix == 0:
            radix = 10
        digits = '0123456789abcdefghijklmnopqrstuvwxyz'[:radix]
        n = 0
        k = 0
        for ch in s.lower():
            if ch not in digits:
                break
            n = n * radix + digits.index(ch)
            k += 1
        if k == 0:
            return math.nan
        return float(sign * n)

    G['parseFloat'] = native('parseFloat', parse_float)
    G['parseInt'] = native('parseInt', parse_int)
    G['isNaN'] = native('isNaN', lambda this, a: tonum(arg(a, 0)) != tonum(arg(a, 0)))
    G['isFinite'] = native('isFinite', lambda this, a: math.isfinite(tonum(arg(a, 0))))
    Number.props.update({
        'parseFloat': G['parseFloat'], 'parseInt': G['parseInt'],
        'isNaN': native('isNaN', lambda this, a: arg(a, 0).__class__ is float and a[0] != a[0]),
        'isFinite': native('isFinite', lambda this, a: arg(a, 0).__class__ is float and math.isfinite(a[0])),
        'isInteger': native('isInteger', lambda this, a: arg(a, 0).__class__ is float and math.isfinite(a[0]) and a[0] == int(a[0])),
        'MAX_VALUE': sys.float_info.max, 'MIN_VALUE': 5e-324, 'EPSILON': sys.float_info.epsilon,
        'MAX_SAFE_INTEGER': 9007199254740991.0, 'MIN_SAFE_INTEGER': -9007199254740991.0,
        'POSITIVE_INFINITY': math.inf, 'NEGATIVE_INFINITY': -math.inf, 'NaN': math.nan,
    })

    def n_to_string(this, args):
        radix = 10 if arg(args, 0) is UNDEF else int(tonum(args[0]))
        if radix == 10:
            return num_to_str(this)
        v = this
        if v != v:
            return 'NaN'
        if v != int(v):
            raise NotImplementedError('fractional toString(radix)')
        n = abs(int(v))
        digits = '0123456789abcdefghijklmnopqrstuvwxyz'
        out = ''
        while True:
            out = digits[n % radix] + out
            n //= radix
            if n == 0:
                break
        return ('-' if v < 0 else '') + out

    method(NumberProto, 'toString', n_to_string)
    method(NumberProto, 'toFixed', lambda this, a: to_fixed(this, 0 if arg(a, 0) is UNDEF else tonum(a[0])))
    method(NumberProto, 'toPrecision', lambda this, a: num_to_str(this) if arg(a, 0) is UNDEF else to_precision(this, tonum(a[0])))
    method(NumberProto, 'valueOf', lambda this, a: this)
    method(BooleanProto, 'toString', lambda this, a: 'true' if this else 'false')
    method(BooleanProto, 'valueOf', lambda this, a: this)

    def regex_of(v):
        if v.__class__ is JSRegExp:
            return v
        return None

    def s_split(this, args):
        sep, lim = arg(args, 0), arg(args, 1)
        if sep is UNDEF:
            parts = [this]
        elif regex_of(sep) is not None:
            parts = [p if p is not None else UNDEF for p in sep.rx.split(this)]
        else:
            sep = tostr(sep)
            parts = list(this) if sep == '' else this.split(sep)
        if lim is not UNDEF:
            parts = parts[:int(tonum(lim))]
        return arr(parts)

    def match_result(m, s):
        r = arr([m.group(0)] + [g if g is not None else UNDEF for g in m.groups()])
        r.props['index'] = float(m.start())
        r.props['input'] = s
        return r

    def s_match(this, args):
        rx = arg(args, 0)
        if regex_of(rx) is None:
            rx = RegExp.ctor([tostr(rx) if rx is not UNDEF else '', ''], RegExp)
        if 'g' in rx.flags:
            found = [m.group(0) for m in rx.rx.finditer(this)]
            return arr(found) if found else None
        m = rx.rx.search(this)
        return match_result(m, this) if m else None

    def expand_replacement(rep, m):
        def sub(mm):
            t = mm.group(1)
            if t == '$':
                return '$'
            if t == '&':
                return m.group(0)
            i = int(t)
            if 1 <= i <= (m.re.groups):
                return m.group(i) or ''
            return mm.group(0)
        return re.sub(r'\$(\$|&|\d{1,2})', sub, rep)

    def s_replace_impl(this, args, all_):
        pat, rep = arg(args, 0), arg(args, 1)

        def repl_for(m):
            if is_callable(rep):
                return tostr(call(rep, UNDEF, [m.group(0)] + [g if g is not None else UNDEF for g in m.groups()] + [float(m.start()), this]))
            return expand_replacement(tostr(rep), m)
        if regex_of(pat) is not None:
            count = 0 if ('g' in pat.flags or all_) else 1
            return pat.rx.sub(repl_for, this, count=count)
        pat = tostr(pat)
        rx = re.compile(re.escape(pat))
        return rx.sub(repl_for, this, count=0 if all_ else 1)

    def s_index_of(this, args):
        return float(this.find(tostr(arg(args, 0)), 0 if arg(args, 1) is UNDEF else int(tonum(args[1]))))

    def s_substring(this, args):
        n = len(this)
        s = 0 if arg(args, 0) is UNDEF else tonum(args[0])
        e = n if arg(args, 1) is UNDEF else tonum(args[1])
        s = int(min(max(0 if s != s else s, 0), n))
        e = int(min(max(0 if e != e else e, 0), n))
        if s > e:
            s, e = e, s
        return this[s:e]

    def s_slice(this, args):
        n = len(this)
        s = to_index(arg(args, 0), n, 0)
        e = to_index(arg(args, 1), n, n)
        return this[s:e] if e > s else ''

    def s_substr(this, args):
        n = len(this)
        s = to_index(arg(args, 0), n, 0)
        cnt = n - s if arg(args, 1) is UNDEF else int(tonum(args[1]))
        return this[s:s + max(cnt, 0)]

    def s_char_code_at(this, args):
        i = 0 if arg(args, 0) is UNDEF else int(tonum(args[0]))
        return float(ord(this[i])) if 0 <= i < len(this) else math.nan

    def s_pad(this, args, left):
        n = int(tonum(arg(args, 0)))
        fill = ' ' if arg(args, 1) is UNDEF else tostr(args[1])
        if n <= len(this) or not fill:
            return this
        pad = (fill * n)[:n - len(this)]
        return pad + this if left else this + pad

    for name, fn in [
        ('split', s_split), ('match', s_match), ('indexOf', s_index_of),
        ('lastIndexOf', lambda this, a: float(this.rfind(tostr(arg(a, 0))))),
        ('replace', lambda this, a: s_replace_impl(this, a, False)),
        ('replaceAll', lambda this, a: s_replace_impl(this, a, True)),
        ('substring', s_substring), ('slice', s_slice), ('substr', s_substr),
        ('trim', lambda this, a: this.strip(' \t\n\r\f\v ﻿')),
        ('trimStart', lambda this, a: this.lstrip(' \t\n\r\f\v ﻿')),
        ('trimEnd', lambda this, a: this.rstrip(' \t\n\r\f\v ﻿')),
        ('toLowerCase', lambda this, a: this.lower()), ('toUpperCase', lambda this, a: this.upper()),
        ('startsWith', lambda this, a: this.startswith(tostr(arg(a, 0)), 0 if arg(a, 1) is UNDEF else int(tonum(a[1])))),
        ('endsWith', lambda this, a: this.endswith(tostr(arg(a, 0)))),
        ('includes', lambda this, a: tostr(arg(a, 0)) in this),
        ('charAt', lambda this, a: this[int(tonum(arg(a, 0)) if arg(a, 0) is not UNDEF else 0)] if 0 <= int(tonum(arg(a, 0)) if arg(a, 0) is not UNDEF else 0) < len(this) else ''),
        ('charCodeAt', s_char_code_at), ('codePointAt', s_char_code_at),
        ('padStart', lambda this, a: s_pad(this, a, True)), ('padEnd', lambda this, a: s_pad(this, a, False)),
        ('repeat', lambda this, a: this * int(tonum(arg(a, 0)))),
        ('concat', lambda this, a: this + ''.join(tostr(x) for x in a)),
        ('toString', lambda this, a: this), ('valueOf', lambda this, a: this),
        ('search', lambda this, a: float(m.start()) if (m := arg(a, 0).rx.search(this)) else -1.0),
    ]:
        method(StringProto, name, fn)

    # ------------------------------------------------------------------------------------------
    # RegExp
    RegExpProto = JSObject(ObjectProto)

    def regexp_ctor(args, nt):
        src, flags = arg(args, 0), arg(args, 1)
        if src.__class__ is JSRegExp:
            flags = src.flags if flags is UNDEF else flags
            src = src.source
        r = JSRegExp(proto_of(nt, RegExpProto))
        r.source = tostr(src)
        r.flags = '' if flags is UNDEF else tostr(flags)
        r.rx = js_regex_to_py(r.source, r.flags)
        r.last_index = 0
        return r

    RegExp = make_ctor('RegExp', RegExpProto, lambda this, a: regexp_ctor(a, RegExp), regexp_ctor)

    def r_exec(this, args):
        s = tostr(arg(args, 0))
        sticky = 'g' in this.flags or 'y' in this.flags
        start = this.last_index if sticky else 0
        m = this.rx.search(s, start) if start <= len(s) else None
        if not m:
            this.last_index = 0
            return None
        if sticky:
            this.last_index = m.end() if m.end() > m.start() else m.end() + 1
        return match_result(m, s)

    method(RegExpProto, 'exec', r_exec)
    method(RegExpProto, 'test', lambda this, a: r_exec(this, a) is not None)
    method(RegExpProto, 'toString', lambda this, a: '/%s/%s' % (this.source, this.flags))

    # ------------------------------------------------------------------------------------------
    # Math
    Math = JSObject(ObjectProto)
    G['Math'] = Math

    def m1(fn):
        def w(this, args):
            x = tonum(arg(args, 0))
            try:
                return float(fn(x))
            except (ValueError, OverflowError):
                return _math_edge(fn, x)
        return w

    def _math_edge(fn, x):
        if x != x:
            return math.nan
        if fn is math.log or fn is math.log2 or fn is math.log10:
            return -math.inf if x == 0 else math.nan
        if fn is math.exp or fn is math.sinh or fn is math.cosh or fn is math.expm1:
            return math.inf if (x > 0 or fn is math.cosh) else (-math.inf if fn is math.sinh else 0.0)
        if fn in (math.floor, math.ceil, math.trunc):
            return x
        if fn is math.log1p:
            return -math.inf if x == -1 else math.nan
        return math.nan

    def m_round(this, args):
        x = tonum(arg(args, 0))
        if x != x or x in (math.inf, -math.inf):
            return x
        r = math.floor(x)
        if x - r >= 0.5:
            r += 1
        if r == 0 and (x < 0 or math.copysign(1.0, x) < 0):
            return -0.0
        return float(r)

    def m_max(this, args):
        r = -math.inf
        for v in args:
            v = v if v.__class__ is float else tonum(v)
            if v != v:
                return math.nan
            if v > r or (v == 0 and r == 0 and math.copysign(1.0, r) < 0):
                r = v
        return r

    def m_min(this, args):
        r = math.inf
        for v in args:
            v = v if v.__class__ is float else tonum(v)
            if v != v:
                return math.nan
            if v < r or (v == 0 and r == 0 and math.copysign(1.0, v) < 0):
                r = v
        return r

    def m_sign(this, args):
        x = tonum(arg(args, 0))
        if x != x or x == 0:
            return x
        return 1.0 if x > 0 else -1.0

    def m_hypot(this, args):
        return math.sqrt(sum(tonum(v) ** 2 for v in args))

    def m_random(this, args):
        if vm.random is None:
            raise RuntimeError('Math.random called but the host installed no generator (vm.random)')
        return vm.random()

    def m_pow(this, args):
        from .interp import js_pow
        return js_pow(tonum(arg(args, 0)), tonum(arg(args, 1)))

    def m_atan2(this, args):
        return math.atan2(tonum(arg(args, 0)), tonum(arg(args, 1)))

    def m_floorlike(fn):
        def w(this, args):
            x = tonum(arg(args, 0))
            if x != x or x in (math.inf, -math.inf):
                return x
            r = float(fn(x))
            if r == 0 and math.copysign(1.0, x) < 0:
                return -0.0
            return r
        return w

    def m_cbrt(x):
        return math.copysign(abs(x) ** (1.0 / 3.0), x)

    Math.props.update({
        'PI': math.pi, 'E': math.e, 'LN2': math.log(2), 'LN10': math.log(10), 'LOG2E': 1 / math.log(2), 'LOG10E': 1 / math.log(10),
        'SQRT2': math.sqrt(2), 'SQRT1_2': math.sqrt(0.5),
        'abs': native('abs', lambda this, a: abs(a[0]) if (a and a[0].__class__ is float) else abs(tonum(arg(a, 0)))),
        'sqrt': native('sqrt', m1(math.sqrt)), 'sin': native('sin', m1(math.sin)), 'cos': native('cos', m1(math.cos)),
        'tan': native('tan', m1(math.tan)), 'asin': native('asin', m1(math.asin)), 'acos': native('acos', m1(math.acos)),
        'atan': native('atan', m1(math.atan)), 'atan2': native('atan2', m_atan2), 'exp': native('exp', m1(math.exp)),
        'log': native('log', m1(math.log)), 'log2': native('log2', m1(math.log2)), 'log10': native('log10', m1(math.log10)),
        'log1p': native('log1p', m1(math.log1p)), 'expm1': native('expm1', m1(math.expm1)),
        'sinh': native('sinh', m1(math.sinh)), 'cosh': native('cosh', m1(math.cosh)), 'tanh': native('tanh', m1(math.tanh)),
        'cbrt': native('cbrt', m1(m_cbrt)),
        'floor': native('floor', m_floorlike(math.floor)), 'ceil': native('ceil', m_floorlike(math.ceil)),
        'trunc': native('trunc', m_floorlike(math.trunc)), 'round': native('round', m_round),
        'max': native('max', m_max), 'min': native('min', m_min), 'sign': native('sign', m_sign), 'hypot': native('hypot', m_hypot),
        'pow': native('pow', m_pow), 'random': native('random', m_random),
        'fround': native('fround', lambda this, a: _fround(tonum(arg(a, 0)))),
    })
    Math.hidden = set(Math.props.keys())

    import array as _array

    def _fround(x):
        return _array.array('f', [x])[0]

    # ------------------------------------------------------------------------------------------
    # JSON
    import json as _json

    def to_json_value(v):
        if v is UNDEF or is_callable(v):
            return _SKIP
        if v is None or v.__class__ in (bool, str):
            return v
        if v.__class__ is float:
            if not math.isfinite(v):
                return None
            return int(v) if v == int(v) and abs(v) < 1e15 else v
        tj = get(v, 'toJSON')
        if is_callable(tj):
            return to_json_value(call(tj, v, []))
        if v.__class__ in (JSArray, ):
            return [None if (x := to_json_value(e)) is _SKIP else x for e in v.items]
        out = {}
        for k in vm.own_keys(v):
            x = to_json_value(get(v, k))
            if x is not _SKIP:
                out[k] = x
        return out

    _SKIP = object()

    def from_json_value(v):
        if isinstance(v, bool) or v is None or isinstance(v, str):
            return v
        if isinstance(v, (int, float)):
            return float(v)
        if isinstance(v, list):
            return arr([from_json_value(x) for x in v])
        o = JSObject(ObjectProto)
        for k, x in v.items():
            o.props[k] = from_json_value(x)
        return o

    def json_stringify(this, args):
        v = to_json_value(arg(args, 0))
        if v is _SKIP:
            return UNDEF
        indent = arg(args, 2)
        if indent is UNDEF:
            return _json.dumps(v, separators=(',', ':'), ensure_ascii=False)
        return _json.dumps(v, indent=int(tonum(indent)) if indent.__class__ is float else tostr(indent), ensure_ascii=False)

    def json_parse(this, args):
        try:
            return from_json_value(_json.loads(tostr(arg(args, 0))))
        except ValueError as e:
            vm.throw('SyntaxError', 'JSON.parse: ' + str(e))

    JSON = JSObject(ObjectProto)
    JSON.props['stringify'] = native('stringify', json_stringify)
    JSON.props['parse'] = native('parse', json_parse)
    G['JSON'] = JSON
    vm.from_py = from_json_value
    vm.to_py = lambda v: None if (x := to_json_value(v)) is _SKIP else x

    # ------------------------------------------------------------------------------------------
    # Date, console
    Date = JSObject(ObjectProto)
    Date.props['now'] = native('now', lambda this, a: float(int(time.time() * 1000)))
    G['Date'] = Date
    console = JSObject(ObjectProto)

    def console_log(this, args):
        sys.stderr.write(' '.join(vm.inspect(a) for a in args) + '\n')
        return UNDEF
    for n in ('log', 'warn', 'error', 'info', 'debug'):
        console.props[n] = native(n, console_log)
    G['console'] = console

    def inspect(v, depth=0):
        if v.__class__ is str:
            return v if depth == 0 else repr(v)
        if not isinstance(v, JSObject):
            return tostr(v)
        if is_callable(v):
            return '[Function %s]' % getattr(v, 'name', '')
        if depth > 2:
            return '[...]'
        if v.__class__ in (JSArray, JSTypedArray):
            return '[' + ', '.join(inspect(x, depth + 1) for x in items_of(v)) + ']'
        return '{' + ', '.join('%s: %s' % (k, inspect(get(v, k), depth + 1)) for k in vm.own_keys(v)) + '}'
    vm.inspect = inspect

    # ------------------------------------------------------------------------------------------
    # Map / Set
    MapProto = JSObject(ObjectProto)
    SetProto = JSObject(ObjectProto)

    def map_ctor(args, nt):
        m = JSMap(proto_of(nt, MapProto))
        m.data = {}
        if arg(args, 0) is not UNDEF and args[0] is not None:
            for e in vm.iterate(args[0]):
                m.data[_key(get(e, 0.0))] = (get(e, 0.0), get(e, 1.0))
        return m

    JSMap.py_iter = lambda self: iter([arr([k, v]) for k, v in list(self.data.values())])
    JSSet.py_iter = lambda self: iter(list(self.data.values()))

    def not_callable(name):
        def f(this, args):
            vm.throw('TypeError', "constructor %s requires 'new'" % name)
        return f

    make_ctor('Map', MapProto, not_callable('Map'), map_ctor)

    def map_set(this, args):
        this.data[_key(arg(args, 0))] = (arg(args, 0), arg(args, 1))
        return this

    def map_for_each(this, args):
        fn = need_fn(arg(args, 0))
        for k, v in list(this.data.values()):
            call(fn, arg(args, 1), [v, k, this])
        return UNDEF

    method(MapProto, 'get', lambda this, a: this.data.get(_key(arg(a, 0)), (UNDEF, UNDEF))[1])
    method(MapProto, 'set', map_set)
    method(MapProto, 'has', lambda this, a: _key(arg(a, 0)) in this.data)
    method(MapProto, 'delete', lambda this, a: this.data.pop(_key(arg(a, 0)), None) is not None)
    method(MapProto, 'clear', lambda this, a: this.data.clear() or UNDEF)
    method(MapProto, 'forEach', map_for_each)
    method(MapProto, 'keys', lambda this, a: list_iter([k for k, _ in this.data.values()]))
    method(MapProto, 'values', lambda this, a: list_iter([v for _, v in this.data.values()]))
    method(MapProto, 'entries', lambda this, a: list_iter([arr([k, v]) for k, v in this.data.values()]))

    def set_ctor(args, nt):
        s = JSSet(proto_of(nt, SetProto))
        s.data = {}
        if arg(args, 0) is not UNDEF and args[0] is not None:
            for e in vm.iterate(args[0]):
                s.data[_key(e)] = e
        return s

    make_ctor('Set', SetProto, not_callable('Set'), set_ctor)

    def set_add(this, args):
        this.data[_key(arg(args, 0))] = arg(args, 0)
        return this

    def set_for_each(this, args):
        fn = need_fn(arg(args, 0))
        for v in list(this.data.values()):
            call(fn, arg(args, 1), [v, v, this])
        return UNDEF

    method(SetProto, 'add', set_add)
    method(SetProto, 'has', lambda this, a: _key(arg(a, 0)) in this.data)
    method(SetProto, 'delete', lambda this, a: this.data.pop(_key(arg(a, 0)), _SKIP) is not _SKIP)
    method(SetProto, 'clear', lambda this, a: this.data.clear() or UNDEF)
    method(SetProto, 'forEach', set_for_each)
    method(SetProto, 'values', lambda this, a: list_iter(list(this.data.values())))
    method(SetProto, 'keys', lambda this, a: list_iter(list(this.data.values())))

    # ------------------------------------------------------------------------------------------
    # generators
    GeneratorProto = JSObject(ObjectProto)
    vm.GeneratorProto = GeneratorProto

    def iter_result(value, done):
        o = JSObject(ObjectProto)
        o.props['value'] = value
        o.props['done'] = done
        return o

    def gen_next(this, args):
        if this.done:
            return iter_result(UNDEF, True)
        try:
            return iter_result(next(this.pygen), False)
        except StopIteration as e:
            this.done = True
            c = e.value
            return iter_result(c.value if c is not None and hasattr(c, 'value') else UNDEF, True)

    method(GeneratorProto, 'next', gen_next)

    def listiter_next(this, args):
        try:
            return iter_result(next(this.it), False)
        except StopIteration:
            return iter_result(UNDEF, True)
    method(ListIterProto, 'next', listiter_next)

    # ------------------------------------------------------------------------------------------
    # Promise (reactions run from vm.run_jobs(), which vm.run() drains before it returns)
    PromiseProto = JSObject(ObjectProto)

    def new_promise(proto=PromiseProto):
        p = JSPromise(proto)
        p.state = 'pending'
        p.value = UNDEF
        p.reactions = []
        return p

    def settle(p, state, value):
        if p.state != 'pending':
            return
        p.state = state
        p.value = value
        if state == 'rejected' and not p.reactions:
            vm.rejections.append(p)       # reported by the host if nothing ever handles it
        for r in p.reactions:
            schedule(p, r)
        p.reactions = []

    def resolve_promise(p, value):
        if p.state != 'pending':
            return
        if value is p:
            return settle(p, 'rejected', 'TypeError: chaining cycle')
        if isinstance(value, JSObject):
            try:
                then = get(value, 'then')
            except JSThrow as t:
                return settle(p, 'rejected', t.value)
            if is_callable(then):
                done = [False]

                def res(this, a):
                    if not done[0]:
                        done[0] = True
                        resolve_promise(p, arg(a, 0))
                    return UNDEF

                def rej(this, a):
                    if not done[0]:
                        done[0] = True
                        settle(p, 'rejected', arg(a, 0))
                    return UNDEF

                def job():
                    try:
                        call(then, value, [native('resolve', res), native('reject', rej)])
                    except JSThrow as t:
                        if not done[0]:
                            done[0] = True
                            settle(p, 'rejected', t.value)
                vm.jobs.append(job)
                return
        settle(p, 'fulfilled', value)

    def schedule(p, reaction):
        on_ok, on_err, child = reaction

        def job():
            handler = on_ok if p.state == 'fulfilled' else on_err
            if not is_callable(handler):
                if p.state == 'fulfilled':
                    resolve_promise(child, p.value)
                else:
                    settle(child, 'rejected', p.value)
                return
            try:
                r = call(handler, UNDEF, [p.value])
            except JSThrow as t:
                settle(child, 'rejected', t.value)
                return
            resolve_promise(child, r)
        vm.jobs.append(job)

    def promise_then(this, args):
        child = new_promise()
        if this in vm.rejections:
            vm.rejections.remove(this)
        r = (arg(args, 0), arg(args, 1), child)
        if this.state == 'pending':
            this.reactions.append(r)
        else:
            schedule(this, r)
        return child

    def promise_ctor(args, nt):
        p = new_promise(proto_of(nt, PromiseProto))
        ex = need_fn(arg(args, 0))
        try:
            call(ex, UNDEF, [native('resolve', lambda this, a: resolve_promise(p, arg(a, 0)) or UNDEF),
                             native('reject', lambda this, a: settle(p, 'rejected', arg(a, 0)) or UNDEF)])
        except JSThrow as t:
            settle(p, 'rejected', t.value)
        return p

    Promise = make_ctor('Promise', PromiseProto, not_callable('Promise'), promise_ctor)
    method(PromiseProto, 'then', promise_then)
    method(PromiseProto, 'catch', lambda this, a: promise_then(this, [UNDEF, arg(a, 0)]))

    def promise_finally(this, args):
        fn = arg(args, 0)

        def ok(t, a):
            call(fn, UNDEF, [])
            return arg(a, 0)

        def bad(t, a):
            call(fn, UNDEF, [])
            raise JSThrow(arg(a, 0))
        return promise_then(this, [native('', ok), native('', bad)])
    method(PromiseProto, 'finally', promise_finally)

    def promise_resolve(this, args):
        v = arg(args, 0)
        if v.__class__ is JSPromise:
            return v
        p = new_promise()
        resolve_promise(p, v)
        return p

    def promise_reject(this, args):
        p = new_promise()
        settle(p, 'rejected', arg(args, 0))
        return p

    def promise_all(this, args):
        items = list(vm.iterate(arg(args, 0)))
        out = new_promise()
        results = [UNDEF] * len(items)
        left = [len(items)]
        if not items:
            resolve_promise(out, arr([]))
        for i, it in enumerate(items):
            def ok(t, a, i=i):
                results[i] = arg(a, 0)
                left[0] -= 1
                if left[0] == 0:
                    resolve_promise(out, arr(results))
                return UNDEF
            promise_then(promise_resolve(None, [it]), [native('', ok), native('', lambda t, a: settle(out, 'rejected', arg(a, 0)) or UNDEF)])
        return out

    Promise.props['resolve'] = native('resolve', promise_resolve)
    Promise.props['reject'] = native('reject', promise_reject)
    Promise.props['all'] = native('all', promise_all)
    vm.promise_resolve = lambda v: promise_resolve(None, [v])
    vm.promise_state = lambda p: (p.state, p.value)

    # Symbol: only what feature tests touch
    Symbol = JSObject(ObjectProto)
    Symbol.props['iterator'] = '@@iterator'
    G['Symbol'] = Symbol

    # eval: global (indirect) evaluation of an expression — what src/serializer.js:109 does with a class name
    def global_eval(this, args):
        src = arg(args, 0)
        if src.__class__ is not str:
            return src
        return vm.eval_expr(src)
    G['eval'] = native('eval', global_eval)

    vm.native = native
    vm.make_ctor = make_ctor
    vm.proto_of = proto_of
    vm.arr = arr
