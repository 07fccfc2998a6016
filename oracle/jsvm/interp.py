"""TEST INFRASTRUCTURE — compiler (AST -> Python closures) and evaluator of `oracle/jsvm`.

See parser.py for what this package is and is not.  `VM` owns one global scope; `vm.run(source)` evaluates a
classic script in it (top-level declarations become globals, which is how the reference's worker loads its
sources: `importScripts(...)`, /root/reference/src/worker.js:3-14).  Built-in objects live in builtins.py.

Statements compile to `exec(scope) -> None | BREAK | CONTINUE | Ret(value)`; expressions to `ev(scope) -> value`.
Generator functions compile their statements a second way, as Python generators, so `yield` can suspend.
"""
import math
import sys

from .parser import parse, Parser
from .runtime import (UNDEF, JSThrow, JSObject, JSArray, JSTypedArray, JSArrayBuffer, JSFunction, NativeFunction, BoundFunction, Scope,
                      typeof, truthy, num_to_str, str_to_num, to_int32, to_uint32)

sys.setrecursionlimit(max(sys.getrecursionlimit(), 12000))

_MISSING = object()


class Ret:
    __slots__ = ('value',)

    def __init__(self, value):
        self.value = value


BREAK = object()
CONTINUE = object()


class JSGenerator(JSObject):
    __slots__ = ('pygen', 'done')


class FnCtx:
    """compile-time context of the function being compiled"""

    def __init__(self, parent, is_arrow, is_gen, strict):
        self.parent = parent
        self.is_arrow = is_arrow
        self.is_gen = is_gen
        self.strict = strict
        self.uses_args = False
        self.uses_super = False
        self.var_names = []

    def nearest_non_arrow(self):
        c = self
        while c is not None and c.is_arrow:
            c = c.parent
        return c


def _contains_yield(node):
    if not isinstance(node, (tuple, list)):
        return False
    if isinstance(node, tuple) and node:
        if node[0] == 'yield':
            return True
        if node[0] in ('function', 'class'):
            return False
    for x in node:
        if isinstance(x, (tuple, list)) and _contains_yield(x):
            return True
    return False


def _collect_var_names(node, out):
    """`var` declarations of a function body (not descending into nested functions)"""
    if not isinstance(node, (tuple, list)):
        return
    if isinstance(node, tuple) and node:
        k = node[0]
        if k in ('function', 'class'):
            return
        if k == 'vardecl' and node[1] == 'var':
            for target, _ in node[2]:
                _pattern_names(target, out)
        if k in ('forof', 'forin') and node[1] == 'var':
            _pattern_names(node[2], out)
    for x in node:
        if isinstance(x, (tuple, list)):
            _collect_var_names(x, out)


def _pattern_names(p, out):
    k = p[0]
    if k == 'id':
        out.append(p[1])
    elif k == 'defpat':
        _pattern_names(p[1], out)
    elif k == 'rest':
        _pattern_names(p[1], out)
    elif k == 'arrpat':
        for e in p[1]:
            if e is not None:
                _pattern_names(e, out)
        if p[2] is not None:
            _pattern_names(p[2], out)
    elif k == 'objpat':
        for _, _, sub in p[1]:
            _pattern_names(sub, out)
        if p[2] is not None:
            _pattern_names(p[2], out)


class VM:
    def __init__(self):
        self.root = Scope(None)
        self.jobs = []          # promise reaction queue
        self.rejections = []    # rejected promises nobody has handled (yet)
        self.random = None      # host-provided Math.random
        from . import builtins
        builtins.install(self)

    # =========================================================================================
    # errors
    def throw(self, kind, msg):
        ctor = self.root.vars.get(kind)
        e = JSObject(ctor.props['prototype'] if ctor is not None else self.ObjectProto)
        e.props['message'] = msg
        e.props['name'] = kind
        raise JSThrow(e)

    # =========================================================================================
    # conversions that may call back into JS
    def toprim(self, v, hint='default'):
        if not isinstance(v, JSObject):
            return v
        order = ('toString', 'valueOf') if hint == 'string' else ('valueOf', 'toString')
        for name in order:
            f = self.get(v, name)
            if isinstance(f, JSObject) and typeof(f) == 'function':
                r = self.call(f, v, [])
                if not isinstance(r, JSObject):
                    return r
        self.throw('TypeError', 'cannot convert object to primitive value')

    def tostr(self, v):
        c = v.__class__
        if c is str:
            return v
        if c is float:
            return num_to_str(v)
        if c is bool:
            return 'true' if v else 'false'
        if v is None:
            return 'null'
        if v is UNDEF:
            return 'undefined'
        return self.tostr(self.toprim(v, 'string'))

    def tonum(self, v):
        c = v.__class__
        if c is float:
            return v
        if c is bool:
            return 1.0 if v else 0.0
        if c is str:
            return str_to_num(v)
        if v is None:
            return 0.0
        if v is UNDEF:
            return math.nan
        return self.tonum(self.toprim(v, 'number'))

    def tokey(self, k):
        c = k.__class__
        if c is str:
            return k
        if c is float:
            return num_to_str(k)
        return self.tostr(k)

    # =========================================================================================
    # property access
    def get(self, o, k):
        c = o.__class__
        if c is JSTypedArray or c is JSArray:
            if k.__class__ is float:
                try:
                    i = int(k)
                except (ValueError, OverflowError):
                    return UNDEF
                items = o.items
                if i == k and 0 <= i < len(items):
                    v = items[i]
                    return float(v) if v.__class__ is int else v
                if c is JSTypedArray or i == k:
                    return UNDEF
                k = num_to_str(k)
            elif k == 'length':
                return float(len(o.items))
            elif k.__class__ is str and k.isdigit():
                i = int(k)
                if i < len(o.items):
                    v = o.items[i]
                    return float(v) if v.__class__ is int else v
                return UNDEF
        elif c is str:
            if k.__class__ is float:
                i = int(k) if k == k and abs(k) < 1e15 else -1
                return o[i] if (i == k and 0 <= i < len(o)) else UNDEF
            if k == 'length':
                return float(len(o))
            o = self.StringProto
        elif c is float:
            o = self.NumberProto
        elif c is bool:
            o = self.BooleanProto
        elif o is None or o is UNDEF:
            self.throw('TypeError', "cannot read properties of %s (reading '%s')" % (self.tostr(o), self.tokey(k)))
        if k.__class__ is not str:
            k = self.tokey(k)
        o0 = o
        while o is not None:
            v = o.props.get(k, _MISSING)
            if v is not _MISSING:
                return v
            o = o.proto
        # computed on demand: typed-array buffer properties, function name / length, Map / Set size
        if c is JSTypedArray:
            if k == 'buffer':
                b = JSArrayBuffer(self.ObjectProto)
                b.owner = o0
                return b
            if k == 'byteOffset':
                return 0.0
            if k == 'byteLength':
                return float(len(o0.items) * o0.items.itemsize)
        elif c is JSArrayBuffer and k == 'byteLength':
            return float(len(o0.owner.items) * o0.owner.items.itemsize)
        if k == 'size' and hasattr(o0, 'data'):
            return float(len(o0.data))
        if (k == 'name' or k == 'length') and c in (JSFunction, NativeFunction, BoundFunction):
            if k == 'name':
                return getattr(o0, 'name', '') or ''
            return getattr(o0, 'nparams', 0.0)
        return UNDEF

    def has(self, o, k):
        if not isinstance(o, JSObject):
            self.throw('TypeError', "cannot use 'in' operator on a primitive")
        c = o.__class__
        if c is JSArray or c is JSTypedArray:
            if k.__class__ is float:
                return k == int(k) and 0 <= k < len(o.items)
            if k == 'length':
                return True
            if k.__class__ is str and k.isdigit():
                return int(k) < len(o.items)
        k = self.tokey(k)
        while o is not None:
            if k in o.props:
                return True
            o = o.proto
        return False

    def put(self, o, k, v):
        c = o.__class__
        if c is JSArray:
            if k.__class__ is float:
                i = int(k) if k == k and abs(k) < 4294967295 else -1
                if i == k and i >= 0:
                    items = o.items
                    n = len(items)
                    if i < n:
                        items[i] = v
                    elif i == n:
                        items.append(v)
                    else:
                        items.extend([UNDEF] * (i - n))
                        items.append(v)
                    return
                k = num_to_str(k)
            elif k == 'length':
                n = int(self.tonum(v))
                items = o.items
                if n < len(items):
                    del items[n:]
                else:
                    items.extend([UNDEF] * (n - len(items)))
                return
            elif k.__class__ is str and k.isdigit():
                return self.put(o, float(k), v)
        elif c is JSTypedArray:
            if k.__class__ is float:
                i = int(k) if k == k and abs(k) < 4294967295 else -1
                if i == k and 0 <= i < len(o.items):
                    if v.__class__ is not float:
                        v = self.tonum(v)
                    o.items[i] = o.conv(v) if o.isint else v
                return
            if k.__class__ is str and k.isdigit():
                return self.put(o, float(k), v)
        elif not isinstance(o, JSObject):
            if o is None or o is UNDEF:
                self.throw('TypeError', "cannot set properties of %s (setting '%s')" % (self.tostr(o), self.tokey(k)))
            return   # assignment to a property of a primitive: silently ignored (sloppy mode)
        if k.__class__ is not str:
            k = self.tokey(k)
        o.props[k] = v

    def delete(self, o, k):
        if not isinstance(o, JSObject):
            return True
        c = o.__class__
        if c is JSArray and k.__class__ is float:
            i = int(k)
            if i == k and 0 <= i < len(o.items):
                o.items[i] = UNDEF
            return True
        o.props.pop(self.tokey(k), None)
        return True

    def own_keys(self, o):
        """enumerable own string keys in JS order: integer keys first, then insertion order"""
        keys = []
        c = o.__class__
        if c is JSArray or c is JSTypedArray:
            keys = [str(i) for i in range(len(o.items))]
        elif c is str:
            return [str(i) for i in range(len(o))]
        if not isinstance(o, JSObject):
            return keys
        hidden = o.hidden
        ints, strs = [], []
        for k in o.props:
            if hidden is not None and k in hidden:
                continue
            if k.isdigit() and (k == '0' or k[0] != '0'):
                ints.append(k)
            else:
                strs.append(k)
        if ints:
            ints.sort(key=int)
        return keys + ints + strs

    # =========================================================================================
    # calls
    def call(self, f, this, args):
        c = f.__class__
        if c is JSFunction:
            return self.call_js(f, this, args, None)
        if c is NativeFunction:
            return f.fn(this, args)
        if c is BoundFunction:
            return self.call(f.target, f.this, f.args + list(args))
        self.throw('TypeError', '%s is not a function' % (self.tostr(f) if not isinstance(f, JSObject) else 'object',))

    def call_js(self, f, this, args, nt):
        if f.is_class and nt is None:
            self.throw('TypeError', "class constructor %s cannot be invoked without 'new'" % f.name)
        scope = Scope(f.env)
        v = scope.vars
        if not f.is_arrow:
            v['this'] = this
            if f.uses_args:
                v['arguments'] = JSArray(self.ArrayProto, list(args))
            if f.home is not None:
                v['%fn'] = f
            if nt is not None:
                v['%nt'] = nt
        for name in f.var_names:
            v[name] = UNDEF
        pn = f.pnames
        if pn is not None:
            n = len(args)
            if n >= len(pn):
                i = 0
                for name in pn:
                    v[name] = args[i]
                    i += 1
            else:
                i = 0
                for name in pn:
                    v[name] = args[i] if i < n else UNDEF
                    i += 1
        else:
            f.binder(scope, args)
        if f.is_gen:
            g = JSGenerator(self.GeneratorProto)
            g.pygen = f.body(scope)
            g.done = False
            return g
        if f.expr_body:
            return f.body(scope)
        c = f.body(scope)
        if c is None:
            return UNDEF
        return c.value

    def construct(self, f, args, nt=None):
        if nt is None:
            nt = f
        c = f.__class__
        if c is NativeFunction:
            if f.ctor is None:
                self.throw('TypeError', '%s is not a constructor' % f.name)
            return f.ctor(args, nt)
        if c is BoundFunction:
            return self.construct(f.target, f.args + list(args), f.target if nt is f else nt)
        if c is not JSFunction or f.is_arrow or (f.home is not None and not f.is_class):
            self.throw('TypeError', 'not a constructor')
        if f.is_class and f.is_derived:
            if f.ctor is None:
                this = self.construct(f.parent, args, nt)
                if f.fields:
                    self.init_fields(f, this)
                return this
            ctor = f.ctor
            scope_this = self.call_ctor(ctor, UNDEF, args, nt)
            return scope_this
        proto = self.get(nt, 'prototype')
        this = JSObject(proto if isinstance(proto, JSObject) else self.ObjectProto)
        if f.is_class:
            if f.fields:
                self.init_fields(f, this)
            if f.ctor is None:
                return this
            return self.call_ctor(f.ctor, this, args, nt)
        return self.call_ctor(f, this, args, nt)

    def call_ctor(self, fn, this, args, nt):
        """run a constructor body; returns the constructed object (`this` after super() for derived classes)"""
        scope = Scope(fn.env)
        v = scope.vars
        v['this'] = this
        v['%fn'] = fn
        v['%nt'] = nt
        if fn.uses_args:
            v['arguments'] = JSArray(self.ArrayProto, list(args))
        for name in fn.var_names:
            v[name] = UNDEF
        if fn.pnames is not None:
            n = len(args)
            i = 0
            for name in fn.pnames:
                v[name] = args[i] if i < n else UNDEF
                i += 1
        else:
            fn.binder(scope, args)
        c = fn.body(scope)
        if c is not None and isinstance(c.value, JSObject):
            return c.value
        this = v['this']
        if this is UNDEF:
            self.throw('ReferenceError', "must call super constructor before accessing 'this'")
        return this

    def init_fields(self, cls, this):
        for key_ev, init_ev, env in cls.fields:
            s = Scope(env)
            s.vars['this'] = this
            k = key_ev(s) if callable(key_ev) else key_ev
            self.put(this, k, init_ev(s) if init_ev is not None else UNDEF)

    def instanceof(self, v, f):
        if not isinstance(f, JSObject):
            self.throw('TypeError', "right-hand side of 'instanceof' is not callable")
        if f.__class__ is BoundFunction:
            f = f.target
        if not isinstance(v, JSObject):
            return False
        proto = self.get(f, 'prototype')
        p = v.proto
        while p is not None:
            if p is proto:
                return True
            p = p.proto
        return False

    # =========================================================================================
    # iteration protocol (host level)
    def iterate(self, v):
        c = v.__class__
        if c is JSArray:
            return _array_iter(v)
        if c is JSTypedArray:
            return iter([float(x) for x in v.items]) if v.isint else iter(list(v.items))
        if c is str:
            return iter(v)
        if c is JSGenerator:
            return v.pygen
        it = getattr(v, 'py_iter', None)
        if it is not None:
            return it()
        self.throw('TypeError', '%s is not iterable' % typeof(v))

    # =========================================================================================
    # operators
    def add(self, a, b):
        ca, cb = a.__class__, b.__class__
        if ca is float and cb is float:
            return a + b
        if ca is str and cb is str:
            return a + b
        a = self.toprim(a)
        b = self.toprim(b)
        if a.__class__ is str or b.__class__ is str:
            return self.tostr(a) + self.tostr(b)
        return self.tonum(a) + self.tonum(b)

    def loose_eq(self, a, b):
        ca, cb = a.__class__, b.__class__
        if ca is cb:
            if ca is float or ca is str:
                return a == b
            return a is b
        if (a is None or a is UNDEF) and (b is None or b is UNDEF):
            return True
        if a is None or a is UNDEF or b is None or b is UNDEF:
            return False
        if ca is bool:
            return self.loose_eq(1.0 if a else 0.0, b)
        if cb is bool:
            return self.loose_eq(a, 1.0 if b else 0.0)
        if ca is float and cb is str:
            return a == str_to_num(b)
        if ca is str and cb is float:
            return str_to_num(a) == b
        if isinstance(a, JSObject) and not isinstance(b, JSObject):
            return self.loose_eq(self.toprim(a), b)
        if isinstance(b, JSObject) and not isinstance(a, JSObject):
            return self.loose_eq(a, self.toprim(b))
        return a is b

    def compare(self, a, b, op):
        """op in '<', '>', '<=', '>=' on arbitrary values"""
        a = self.toprim(a, 'number')
        b = self.toprim(b, 'number')
        if a.__class__ is str and b.__class__ is str:
            pass
        else:
            a = self.tonum(a)
            b = self.tonum(b)
        if op == '<':
            return a < b
        if op == '>':
            return a > b
        if op == '<=':
            return a <= b
        return a >= b

    # =========================================================================================
    # running code
    def run(self, src, filename='<js>'):
        ast = parse(src, filename)
        fctx = FnCtx(None, False, False, False)
        names = []
        _collect_var_names(ast[1], names)
        for n in names:
            self.root.vars.setdefault(n, UNDEF)
        ex = self.c_block_body(ast[1], fctx, new_scope=False)
        self.root.vars.setdefault('this', self.global_object)
        c = ex(self.root)
        self.run_jobs()
        return c.value if isinstance(c, Ret) else UNDEF

    def eval_expr(self, src):
        e = Parser(src).parse_expression_only()
        return self.c_expr(e, FnCtx(None, False, False, False))(self.root)

    def run_jobs(self):
        while self.jobs:
            job = self.jobs.pop(0)
            job()

    def make_function_from_source(self, params_src, body_src):
        """`new Function(...)`: sloppy-mode function in the global scope"""
        p = Parser('(function anonymous(%s\n) {\n%s\n})' % (params_src, body_src), '<Function>')
        e = p.parse_expression_only()
        return self.c_expr(e, FnCtx(None, False, False, False))(self.root)

    # =========================================================================================
    # compiler: statements
    def c_block_body(self, stmts, fctx, new_scope):
        """compile a statement list; function declarations are hoisted to the top of the block"""
        hoisted = [self.c_stmt(s, fctx) for s in stmts if s[0] == 'funcdecl']
        rest = [self.c_stmt(s, fctx) for s in stmts if s[0] != 'funcdecl']
        seq = tuple(hoisted + rest)
        needs_scope = new_scope and any(s[0] in ('funcdecl', 'classdecl') or (s[0] == 'vardecl' and s[1] != 'var') for s in stmts)
        if len(seq) == 1 and not needs_scope:
            return seq[0]
        if needs_scope:
            def ex_scoped(env):
                env = Scope(env)
                for s in seq:
                    c = s(env)
                    if c is not None:
                        return c
                return None
            return ex_scoped

        def ex(env):
            for s in seq:
                c = s(env)
                if c is not None:
                    return c
            return None
        return ex

    def c_stmt(self, node, fctx):
        k = node[0]
        m = getattr(self, 's_' + k, None)
        if m is None:
            raise NotImplementedError('statement ' + k)
        return m(node, fctx)

    def s_empty(self, node, fctx):
        return lambda env: None

    def s_expr(self, node, fctx):
        e = self.c_expr(node[1], fctx)

        def ex(env):
            e(env)
            return None
        return ex

    def s_block(self, node, fctx):
        return self.c_block_body(node[1], fctx, new_scope=True)

    def s_vardecl(self, node, fctx):
        kind = node[1]
        parts = []
        for target, init in node[2]:
            init_ev = self.c_expr(init, fctx, name_hint=target[1] if target[0] == 'id' else None) if init is not None else None
            if kind == 'var':
                if init_ev is None:
                    continue
                parts.append((self.c_assign_target(target, fctx), init_ev))
            else:
                parts.append((self.c_declare(target, fctx), init_ev))
        if len(parts) == 1:
            b, i = parts[0]
            if i is None:
                def ex1n(env):
                    b(env, UNDEF)
                    return None
                return ex1n

            def ex1(env):
                b(env, i(env))
                return None
            return ex1

        def ex(env):
            for b, i in parts:
                b(env, i(env) if i is not None else UNDEF)
            return None
        return ex

    def s_funcdecl(self, node, fctx):
        mk = self.c_function(node[2], fctx)
        name = node[1]

        def ex(env):
            env.vars[name] = mk(env)
            return None
        return ex

    def s_classdecl(self, node, fctx):
        mk = self.c_class(node[2], fctx)
        name = node[1]

        def ex(env):
            env.vars[name] = mk(env)
            return None
        return ex

    def s_return(self, node, fctx):
        if node[1] is None:
            r = Ret(UNDEF)
            return lambda env: r
        e = self.c_expr(node[1], fctx)
        return lambda env: Ret(e(env))

    def s_if(self, node, fctx):
        test = self.c_cond(node[1], fctx)
        cons = self.c_stmt(node[2], fctx)
        if node[3] is None:
            def ex(env):
                if test(env):
                    return cons(env)
                return None
            return ex
        alt = self.c_stmt(node[3], fctx)

        def ex2(env):
            if test(env):
                return cons(env)
            return alt(env)
        return ex2

    def s_for(self, node, fctx):
        _, init, test, update, body = node
        init_ex = self.c_stmt(init, fctx) if init is not None else None
        test_ev = self.c_cond(test, fctx) if test is not None else None
        upd_ev = self.c_expr(update, fctx) if update is not None else None
        body_ex = self.c_stmt(body, fctx)
        lexical = init is not None and init[0] == 'vardecl' and init[1] != 'var'
        # a fresh binding per iteration is only observable when a closure created in the body outlives it
        per_iter = lexical and _contains_function(body)

        def ex(env):
            if lexical:
                env = Scope(env)
            if init_ex is not None:
                init_ex(env)
            while True:
                if test_ev is not None and not test_ev(env):
                    break
                c = body_ex(env)
                if c is not None:
                    if c is BREAK:
                        break
                    if c is not CONTINUE:
                        return c
                if per_iter:
                    nxt = Scope(env.parent)
                    nxt.vars.update(env.vars)
                    env = nxt
                if upd_ev is not None:
                    upd_ev(env)
            return None
        return ex

    def s_forof(self, node, fctx):
        _, kind, target, it, body = node
        it_ev = self.c_expr(it, fctx)
        body_ex = self.c_stmt(body, fctx)
        if kind is None or kind == 'var':
            bind = self.c_assign_target(target, fctx)
            fresh = False
        else:
            bind = self.c_declare(target, fctx)
            fresh = True
        iterate = self.iterate

        def ex(env):
            for v in iterate(it_ev(env)):
                s = Scope(env) if fresh else env
                bind(s, v)
                c = body_ex(s)
                if c is not None:
                    if c is BREAK:
                        break
                    if c is not CONTINUE:
                        return c
            return None
        return ex

    def s_forin(self, node, fctx):
        _, kind, target, obj, body = node
        obj_ev = self.c_expr(obj, fctx)
        body_ex = self.c_stmt(body, fctx)
        if kind is None or kind == 'var':
            bind = self.c_assign_target(target, fctx)
            fresh = False
        else:
            bind = self.c_declare(target, fctx)
            fresh = True

        def ex(env):
            o = obj_ev(env)
            if o is None or o is UNDEF:
                return None
            keys, seen = [], set()
            p = o
            while p is not None and isinstance(p, (JSObject, str)):
                for k in self.own_keys(p):
                    if k not in seen:
                        seen.add(k)
                        keys.append(k)
                p = p.proto if isinstance(p, JSObject) else None
            for k in keys:
                s = Scope(env) if fresh else env
                bind(s, k)
                c = body_ex(s)
                if c is not None:
                    if c is BREAK:
                        break
                    if c is not CONTINUE:
                        return c
            return None
        return ex

    def s_while(self, node, fctx):
        test = self.c_cond(node[1], fctx)
        body = self.c_stmt(node[2], fctx)

        def ex(env):
            while test(env):
                c = body(env)
                if c is not None:
                    if c is BREAK:
                        break
                    if c is not CONTINUE:
                        return c
            return None
        return ex

    def s_dowhile(self, node, fctx):
        body = self.c_stmt(node[1], fctx)
        test = self.c_cond(node[2], fctx)

        def ex(env):
            while True:
                c = body(env)
                if c is not None:
                    if c is BREAK:
                        break
                    if c is not CONTINUE:
                        return c
                if not test(env):
                    break
            return None
        return ex

    def s_break(self, node, fctx):
        return lambda env: BREAK

    def s_continue(self, node, fctx):
        return lambda env: CONTINUE

    def s_throw(self, node, fctx):
        e = self.c_expr(node[1], fctx)

        def ex(env):
            raise JSThrow(e(env))
        return ex

    def s_try(self, node, fctx):
        _, blk, param, handler, final = node
        blk_ex = self.c_stmt(blk, fctx)
        h_ex = self.c_stmt(handler, fctx) if handler is not None else None
        f_ex = self.c_stmt(final, fctx) if final is not None else None
        p_bind = self.c_declare(param, fctx) if param is not None else None

        def ex(env):
            try:
                try:
                    c = blk_ex(env)
                except JSThrow as t:
                    if h_ex is None:
                        raise
                    s = Scope(env)
                    if p_bind is not None:
                        p_bind(s, t.value)
                    c = h_ex(s)
                except RecursionError:
                    if h_ex is None:
                        raise
                    s = Scope(env)
                    if p_bind is not None:
                        p_bind(s, 'RangeError: Maximum call stack size exceeded')
                    c = h_ex(s)
            finally:
                if f_ex is not None:
                    fc = f_ex(env)
                    if fc is not None:
                        return fc
            return c
        return ex

    def s_switch(self, node, fctx):
        disc = self.c_expr(node[1], fctx)
        cases = [(self.c_expr(t, fctx) if t is not None else None, [self.c_stmt(s, fctx) for s in body]) for t, body in node[2]]

        def ex(env):
            env = Scope(env)
            d = disc(env)
            start = None
            for idx, (t, _) in enumerate(cases):
                if t is not None and _strict_eq(d, t(env)):
                    start = idx
                    break
            if start is None:
                for idx, (t, _) in enumerate(cases):
                    if t is None:
                        start = idx
                        break
            if start is None:
                return None
            for _, body in cases[start:]:
                for s in body:
                    c = s(env)
                    if c is not None:
                        if c is BREAK:
                            return None
                        return c
            return None
        return ex

    # =========================================================================================
    # compiler: generator-mode statements (Python generators that yield JS values and return a completion)
    def g_block_body(self, stmts, fctx, new_scope):
        parts = [self.g_stmt(s, fctx) for s in stmts if s[0] == 'funcdecl'] + [self.g_stmt(s, fctx) for s in stmts if s[0] != 'funcdecl']
        needs_scope = new_scope and any(s[0] in ('funcdecl', 'classdecl') or (s[0] == 'vardecl' and s[1] != 'var') for s in stmts)

        def gen(env):
            if needs_scope:
                env = Scope(env)
            for is_gen, s in parts:
                if is_gen:
                    c = yield from s(env)
                else:
                    c = s(env)
                if c is not None:
                    return c
            return None
        return gen

    def g_stmt(self, node, fctx):
        """returns (is_generator, exec)"""
        if not _contains_yield(node):
            return (False, self.c_stmt(node, fctx))
        k = node[0]
        if k == 'expr' and node[1][0] == 'yield':
            _, arg, delegate = node[1]
            arg_ev = self.c_expr(arg, fctx) if arg is not None else None
            iterate = self.iterate

            def g_yield(env):
                v = arg_ev(env) if arg_ev is not None else UNDEF
                if delegate:
                    for x in iterate(v):
                        yield x
                else:
                    yield v
                return None
            return (True, g_yield)
        if k == 'block':
            return (True, self.g_block_body(node[1], fctx, True))
        if k == 'if':
            test = self.c_cond(node[1], fctx)
            cg, cons = self.g_stmt(node[2], fctx)
            ag, alt = self.g_stmt(node[3], fctx) if node[3] is not None else (False, None)

            def g_if(env):
                if test(env):
                    return (yield from cons(env)) if cg else cons(env)
                if alt is not None:
                    return (yield from alt(env)) if ag else alt(env)
                return None
            return (True, g_if)
        if k == 'for':
            _, init, test, update, body = node
            init_ex = self.c_stmt(init, fctx) if init is not None else None
            test_ev = self.c_cond(test, fctx) if test is not None else None
            upd_ev = self.c_expr(update, fctx) if update is not None else None
            _, body_g = self.g_stmt(body, fctx)

            def g_for(env):
                env = Scope(env)
                if init_ex is not None:
                    init_ex(env)
                while True:
                    if test_ev is not None and not test_ev(env):
                        break
                    c = yield from body_g(env)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
                    if upd_ev is not None:
                        upd_ev(env)
                return None
            return (True, g_for)
        if k == 'forof':
            _, kind, target, it, body = node
            it_ev = self.c_expr(it, fctx)
            _, body_g = self.g_stmt(body, fctx)
            bind = self.c_assign_target(target, fctx) if kind in (None, 'var') else self.c_declare(target, fctx)
            iterate = self.iterate

            def g_forof(env):
                for v in iterate(it_ev(env)):
                    s = Scope(env)
                    bind(s, v)
                    c = yield from body_g(s)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
                return None
            return (True, g_forof)
        if k == 'while':
            test = self.c_cond(node[1], fctx)
            _, body_g = self.g_stmt(node[2], fctx)

            def g_while(env):
                while test(env):
                    c = yield from body_g(env)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
                return None
            return (True, g_while)
        raise NotImplementedError('yield inside a %s statement (only yield statements in blocks / if / for / for-of / while)' % k)

    # =========================================================================================
    # compiler: binding patterns
    def c_declare(self, p, fctx):
        """binder(env, value): declare the pattern's names in env (let / const / parameters / catch)"""
        k = p[0]
        if k == 'id':
            name = p[1]

            def b_id(env, v):
                env.vars[name] = v
            return b_id
        return self.c_pattern(p, fctx, self.c_declare)

    def c_assign_target(self, p, fctx):
        """binder(env, value): assign to existing bindings / properties"""
        k = p[0]
        if k == 'id':
            name = p[1]
            root = self.root

            def a_id(env, v):
                s = env
                while s is not None:
                    if name in s.vars:
                        s.vars[name] = v
                        return
                    s = s.parent
                root.vars[name] = v      # sloppy mode: implicit global
            return a_id
        if k == 'member':
            obj = self.c_expr(p[1], fctx)
            name = p[2]
            put = self.put

            def a_member(env, v):
                put(obj(env), name, v)
            return a_member
        if k == 'index':
            obj = self.c_expr(p[1], fctx)
            idx = self.c_expr(p[2], fctx)
            put = self.put

            def a_index(env, v):
                o = obj(env)
                put(o, idx(env), v)
            return a_index
        return self.c_pattern(p, fctx, self.c_assign_target)

    def c_pattern(self, p, fctx, leaf):
        k = p[0]
        if k == 'defpat':
            inner = leaf(p[1], fctx)
            default = self.c_expr(p[2], fctx)

            def b_def(env, v):
                inner(env, default(env) if v is UNDEF else v)
            return b_def
        if k == 'arrpat':
            elems = [leaf(e, fctx) if e is not None else None for e in p[1]]
            rest = leaf(p[2], fctx) if p[2] is not None else None
            iterate = self.iterate
            ArrayProto = self.ArrayProto

            def b_arr(env, v):
                c = v.__class__
                if c is JSArray or c is JSTypedArray:
                    items = v.items
                    n = len(items)
                    i = 0
                    for e in elems:
                        if e is not None:
                            if i < n:
                                x = items[i]
                                if x.__class__ is int:
                                    x = float(x)
                            else:
                                x = UNDEF
                            e(env, x)
                        i += 1
                    if rest is not None:
                        rest(env, JSArray(ArrayProto, [float(x) if x.__class__ is int else x for x in items[i:]]))
                    return
                it = iterate(v)
                for e in elems:
                    x = next(it, UNDEF)
                    if e is not None:
                        e(env, x)
                if rest is not None:
                    rest(env, JSArray(ArrayProto, list(it)))
            return b_arr
        if k == 'objpat':
            props = []
            for key, computed, sub in p[1]:
                props.append((self.c_expr(key, fctx) if computed else key, computed, leaf(sub, fctx)))
            rest = leaf(p[2], fctx) if p[2] is not None else None
            get = self.get

            def b_obj(env, v):
                if v is None or v is UNDEF:
                    self.throw('TypeError', 'cannot destructure %s' % self.tostr(v))
                used = []
                for key, computed, sub in props:
                    kk = key(env) if computed else key
                    used.append(kk)
                    sub(env, get(v, kk))
                if rest is not None:
                    o = JSObject(self.ObjectProto)
                    for kk in self.own_keys(v):
                        if kk not in used:
                            o.props[kk] = get(v, kk)
                    rest(env, o)
            return b_obj
        raise NotImplementedError('pattern ' + k)

    # =========================================================================================
    # compiler: functions and classes
    def c_function(self, node, fctx, name_hint=None, home_kind=None, strict=None):
        """returns mk(env) -> JSFunction"""
        _, name, params, body, flags = node
        is_arrow = flags['arrow']
        is_gen = flags['gen']
        is_method = flags.get('method', False)
        expr_body = flags['expr_body']
        if strict is None:
            strict = fctx.strict
        inner = FnCtx(fctx, is_arrow, is_gen, strict)
        # parameters
        simple = all(p[0] == 'id' for p in params)
        pnames = tuple(p[1] for p in params) if simple else None
        binder = None
        if not simple:
            binders = []
            for p in params:
                if p[0] == 'rest':
                    binders.append(('rest', self.c_declare(p[1], inner)))
                else:
                    binders.append(('one', self.c_declare(p, inner)))
            ArrayProto = self.ArrayProto

            def binder(scope, args):
                n = len(args)
                i = 0
                for kind, b in binders:
                    if kind == 'one':
                        b(scope, args[i] if i < n else UNDEF)
                        i += 1
                    else:
                        b(scope, JSArray(ArrayProto, list(args[i:])))
        if expr_body:
            body_c = self.c_expr(body, inner)
            var_names = ()
        else:
            names = []
            _collect_var_names(body[1], names)
            pset = set(pnames or ())
            var_names = tuple(n for n in dict.fromkeys(names) if n not in pset)
            if is_gen:
                body_c = self.g_block_body(body[1], inner, new_scope=False)
            else:
                body_c = self.c_block_body(body[1], inner, new_scope=False)
        fname = name or name_hint or ''
        vm = self
        FunctionProto = self.FunctionProto
        ObjectProto = self.ObjectProto
        plain = not is_arrow and not is_method and not is_gen
        sloppy_this = plain and not strict
        nparams = float(len([p for p in params if p[0] == 'id']))

        if sloppy_this:
            # sloppy-mode functions see the global object as `this` when called without a receiver
            raw_body = body_c
            glob = self

            def body_sloppy(scope):
                t = scope.vars.get('this')
                if t is UNDEF or t is None:
                    scope.vars['this'] = glob.global_object
                return raw_body(scope)
            body_c = body_sloppy

        def mk(env):
            f = JSFunction.__new__(JSFunction)
            f.props = {}
            f.proto = FunctionProto
            f.hidden = None
            f.cls = 'Function'
            f.name = fname
            f.env = env
            f.pnames = pnames
            f.binder = binder
            f.body = body_c
            f.is_arrow = is_arrow
            f.is_gen = is_gen
            f.expr_body = expr_body
            f.home = None
            f.is_class = False
            f.is_derived = False
            f.parent = None
            f.ctor = None
            f.fields = None
            f.uses_args = inner.uses_args
            f.nparams = nparams
            f.vm = vm
            f.var_names = var_names
            if plain:
                po = JSObject(ObjectProto)
                po.props['constructor'] = f
                po.hidden = {'constructor'}
                f.props['prototype'] = po
            if name and not is_method and not is_arrow:
                # a named function (expression or declaration) sees its own name
                s = Scope(env)
                s.vars[name] = f
                f.env = s
            return f
        return mk

    def c_class(self, node, fctx, name_hint=None):
        _, name, parent, ctor, methods, fields = node
        cname = name or name_hint or ''
        parent_ev = self.c_expr(parent, fctx) if parent is not None else None
        cctx = FnCtx(fctx, True, False, True)   # class bodies are strict; `this` in field initialisers is bound per instance
        ctor_mk = self.c_function(ctor, cctx, strict=True) if ctor is not None else None
        meths = []
        for static, key, computed, fn in methods:
            meths.append((static, self.c_expr(key, fctx) if computed else key, computed, self.c_function(fn, cctx, strict=True)))
        flds = []
        for static, key, computed, init in fields:
            ictx = FnCtx(fctx, False, False, True)
            flds.append((static, self.c_expr(key, fctx) if computed else key, computed,
                         self.c_expr(init, ictx) if init is not None else None))
        vm = self

        def mk(env):
            par = UNDEF
            if parent_ev is not None:
                par = parent_ev(env)
                if par is not None and typeof(par) != 'function':
                    vm.throw('TypeError', 'class extends value is not a constructor or null')
            cenv = Scope(env)
            F = JSFunction.__new__(JSFunction)
            F.props = {}
            F.hidden = {'prototype'}
            F.cls = 'Function'
            F.name = cname
            F.env = cenv
            F.pnames = ()
            F.binder = None
            F.body = None
            F.is_arrow = False
            F.is_gen = False
            F.expr_body = False
            F.is_class = True
            F.is_derived = parent_ev is not None
            F.parent = par if parent_ev is not None else None
            F.uses_args = False
            F.nparams = 0.0
            F.vm = vm
            F.var_names = ()
            F.home = None
            F.ctor = None
            F.fields = None
            if parent_ev is not None and par is not None:
                F.proto = par
                pp = vm.get(par, 'prototype')
                protoobj = JSObject(pp if isinstance(pp, JSObject) else None)
            else:
                F.proto = vm.FunctionProto
                protoobj = JSObject(vm.ObjectProto if parent_ev is None else None)
            protoobj.props['constructor'] = F
            protoobj.hidden = {'constructor'}
            F.props['prototype'] = protoobj
            if name:
                cenv.vars[name] = F
            if ctor_mk is not None:
                c = ctor_mk(cenv)
                c.home = protoobj
                c.props.pop('prototype', None)
                F.ctor = c
                F.nparams = c.nparams
            else:
                F.ctor = None
            for static, key, computed, fmk in meths:
                f = fmk(cenv)
                target = F if static else protoobj
                f.home = target
                k = key(cenv) if computed else key
                f.name = k if isinstance(k, str) else ''
                target.props[vm.tokey(k)] = f
                if target.hidden is None:
                    target.hidden = set()
                target.hidden.add(vm.tokey(k))
            inst = []
            F.fields = inst
            for static, key, computed, init in flds:
                k = key(cenv) if computed else key
                if static:
                    s = Scope(cenv)
                    s.vars['this'] = F
                    vm.put(F, k, init(s) if init is not None else UNDEF)
                else:
                    inst.append((k, init, cenv))
            return F
        return mk

    # =========================================================================================
    # compiler: expressions
    def c_cond(self, node, fctx):
        """expression compiled for its truth value (returns a Python bool)"""
        k = node[0]
        if k == 'paren':
            return self.c_cond(node[1], fctx)
        if k == 'bin' and node[1] in ('<', '>', '<=', '>=', '==', '!=', '===', '!=='):
            return self.c_expr(node, fctx)
        if k == 'unary' and node[1] == '!':
            inner = self.c_cond(node[2], fctx)
            return lambda env: not inner(env)
        if k == 'logical' and node[1] in ('&&', '||'):
            a = self.c_cond(node[2], fctx)
            b = self.c_cond(node[3], fctx)
            if node[1] == '&&':
                return lambda env: a(env) and b(env)
            return lambda env: a(env) or b(env)
        e = self.c_expr(node, fctx)
        return lambda env: truthy(e(env))

    def c_expr(self, node, fctx, name_hint=None):
        k = node[0]
        if k == 'function':
            return self.c_function(node, fctx, name_hint=name_hint)
        if k == 'class':
            return self.c_class(node, fctx, name_hint=name_hint)
        m = getattr(self, 'e_' + k, None)
        if m is None:
            raise NotImplementedError('expression ' + k)
        return m(node, fctx)

    def e_num(self, node, fctx):
        v = node[1]
        return lambda env: v

    e_str = e_num
    e_bool = e_num

    def e_null(self, node, fctx):
        return lambda env: None

    def e_paren(self, node, fctx):
        return self.c_expr(node[1], fctx)

    def e_regex(self, node, fctx):
        ctor = self.root.vars['RegExp']
        body, flags = node[1], node[2]
        return lambda env: ctor.ctor([body, flags], ctor)

    def e_template(self, node, fctx):
        strs = node[1]
        exprs = [self.c_expr(e, fctx) for e in node[2]]
        tostr = self.tostr

        def ev(env):
            out = [strs[0]]
            for i, e in enumerate(exprs):
                out.append(tostr(e(env)))
                out.append(strs[i + 1])
            return ''.join(out)
        return ev

    def e_id(self, node, fctx):
        name = node[1]
        if name == 'undefined':
            return lambda env: UNDEF
        if name == 'arguments':
            c = fctx.nearest_non_arrow()
            if c is not None:
                c.uses_args = True
        vm = self

        def ev(env):
            s = env
            while s is not None:
                v = s.vars
                if name in v:
                    return v[name]
                s = s.parent
            vm.throw('ReferenceError', name + ' is not defined')
        return ev

    def e_this(self, node, fctx):
        def ev(env):
            s = env
            while s is not None:
                v = s.vars
                if 'this' in v:
                    return v['this']
                s = s.parent
            return UNDEF
        return ev

    def e_array(self, node, fctx):
        elems = []
        spread = False
        for e in node[1]:
            if e is None:
                elems.append((0, None))
            elif e[0] == 'spread':
                spread = True
                elems.append((2, self.c_expr(e[1], fctx)))
            else:
                elems.append((1, self.c_expr(e, fctx)))
        ArrayProto = self.ArrayProto
        if not spread and all(kind == 1 for kind, _ in elems):
            evs = tuple(e for _, e in elems)
            return lambda env: JSArray(ArrayProto, [e(env) for e in evs])
        iterate = self.iterate

        def ev(env):
            out = []
            for kind, e in elems:
                if kind == 1:
                    out.append(e(env))
                elif kind == 2:
                    out.extend(iterate(e(env)))
                else:
                    out.append(UNDEF)
            return JSArray(ArrayProto, out)
        return ev

    def e_object(self, node, fctx):
        parts = []
        for p in node[1]:
            if p[0] == 'spread':
                parts.append((2, None, self.c_expr(p[1], fctx)))
            elif p[0] == 'shorthand_default':
                raise NotImplementedError('shorthand default outside a pattern')
            else:
                _, key, computed, val = p
                is_method = val[0] == 'function' and val[4].get('method')
                if val[0] == 'function':
                    vev = self.c_function(val, fctx, name_hint=key if not computed else None)
                else:
                    vev = self.c_expr(val, fctx, name_hint=key if not computed else None)
                parts.append((3 if is_method else (1 if computed else 0), self.c_expr(key, fctx) if computed else key, vev))
        ObjectProto = self.ObjectProto
        vm = self
        if all(kind == 0 for kind, _, _ in parts):
            kv = tuple((k, v) for _, k, v in parts)

            def ev_simple(env):
                o = JSObject(ObjectProto)
                props = o.props
                for k, v in kv:
                    props[k] = v(env)
                return o
            return ev_simple

        def ev(env):
            o = JSObject(ObjectProto)
            for kind, k, v in parts:
                if kind == 0:
                    o.props[k] = v(env)
                elif kind == 1:
                    o.props[vm.tokey(k(env))] = v(env)
                elif kind == 3:
                    f = v(env)
                    f.home = o
                    o.props[k if isinstance(k, str) else vm.tokey(k(env))] = f
                else:
                    src = v(env)
                    if isinstance(src, (JSObject, str)):
                        for kk in vm.own_keys(src):
                            o.props[kk] = vm.get(src, kk)
            return o
        return ev

    def e_seq(self, node, fctx):
        es = [self.c_expr(e, fctx) for e in node[1]]

        def ev(env):
            v = UNDEF
            for e in es:
                v = e(env)
            return v
        return ev

    def e_cond(self, node, fctx):
        t = self.c_cond(node[1], fctx)
        a = self.c_expr(node[2], fctx)
        b = self.c_expr(node[3], fctx)
        return lambda env: a(env) if t(env) else b(env)

    def e_logical(self, node, fctx):
        op = node[1]
        a = self.c_expr(node[2], fctx)
        b = self.c_expr(node[3], fctx)
        if op == '&&':
            def ev_and(env):
                v = a(env)
                return b(env) if truthy(v) else v
            return ev_and
        if op == '||':
            def ev_or(env):
                v = a(env)
                return v if truthy(v) else b(env)
            return ev_or

        def ev_nullish(env):
            v = a(env)
            return b(env) if (v is None or v is UNDEF) else v
        return ev_nullish

    def e_unary(self, node, fctx):
        op = node[1]
        if op == 'typeof':
            if node[2][0] == 'id':
                name = node[2][1]

                def ev_typeof_id(env):
                    s = env
                    while s is not None:
                        if name in s.vars:
                            return typeof(s.vars[name])
                        s = s.parent
                    return 'undefined'
                return ev_typeof_id
            a = self.c_expr(node[2], fctx)
            return lambda env: typeof(a(env))
        if op == 'delete':
            t = node[2]
            if t[0] == 'member':
                obj = self.c_expr(t[1], fctx)
                name = t[2]
                return lambda env: self.delete(obj(env), name)
            if t[0] == 'index':
                obj = self.c_expr(t[1], fctx)
                idx = self.c_expr(t[2], fctx)
                return lambda env: self.delete(obj(env), idx(env))
            return lambda env: True
        a = self.c_expr(node[2], fctx)
        tonum = self.tonum
        if op == '!':
            return lambda env: not truthy(a(env))
        if op == '-':
            def ev_neg(env):
                v = a(env)
                return -v if v.__class__ is float else -tonum(v)
            return ev_neg
        if op == '+':
            return lambda env: tonum(a(env))
        if op == '~':
            return lambda env: float(~to_int32(tonum(a(env))))
        if op == 'void':
            def ev_void(env):
                a(env)
                return UNDEF
            return ev_void
        raise NotImplementedError(op)

    def e_bin(self, node, fctx):
        op = node[1]
        a = self.c_expr(node[2], fctx)
        b = self.c_expr(node[3], fctx)
        vm = self
        tonum = self.tonum
        if op == '+':
            add = self.add

            def ev_add(env):
                x = a(env)
                y = b(env)
                if x.__class__ is float and y.__class__ is float:
                    return x + y
                return add(x, y)
            return ev_add
        if op == '-':
            def ev_sub(env):
                x = a(env)
                y = b(env)
                if x.__class__ is float and y.__class__ is float:
                    return x - y
                return tonum(x) - tonum(y)
            return ev_sub
        if op == '*':
            def ev_mul(env):
                x = a(env)
                y = b(env)
                if x.__class__ is float and y.__class__ is float:
                    return x * y
                return tonum(x) * tonum(y)
            return ev_mul
        if op == '/':
            def ev_div(env):
                x = a(env)
                y = b(env)
                if x.__class__ is not float:
                    x = tonum(x)
                if y.__class__ is not float:
                    y = tonum(y)
                try:
                    return x / y
                except ZeroDivisionError:
                    if x != x or x == 0:
                        return math.nan
                    return math.inf if (x > 0) == (math.copysign(1.0, y) > 0) else -math.inf
            return ev_div
        if op == '%':
            def ev_mod(env):
                x = tonum(a(env))
                y = tonum(b(env))
                if y == 0 or x != x or y != y or x in (math.inf, -math.inf):
                    return math.nan
                if y in (math.inf, -math.inf):
                    return x
                return math.fmod(x, y)
            return ev_mod
        if op == '**':
            return lambda env: js_pow(tonum(a(env)), tonum(b(env)))
        if op in ('<', '>', '<=', '>='):
            compare = self.compare
            if op == '<':
                def ev_lt(env):
                    x = a(env)
                    y = b(env)
                    if x.__class__ is float and y.__class__ is float:
                        return x < y
                    return compare(x, y, '<')
                return ev_lt
            if op == '>':
                def ev_gt(env):
                    x = a(env)
                    y = b(env)
                    if x.__class__ is float and y.__class__ is float:
                        return x > y
                    return compare(x, y, '>')
                return ev_gt
            if op == '<=':
                def ev_le(env):
                    x = a(env)
                    y = b(env)
                    if x.__class__ is float and y.__class__ is float:
                        return x <= y
                    return compare(x, y, '<=')
                return ev_le

            def ev_ge(env):
                x = a(env)
                y = b(env)
                if x.__class__ is float and y.__class__ is float:
                    return x >= y
                return compare(x, y, '>=')
            return ev_ge
        if op == '===':
            return lambda env: _strict_eq(a(env), b(env))
        if op == '!==':
            return lambda env: not _strict_eq(a(env), b(env))
        if op == '==':
            leq = self.loose_eq

            def ev_eq(env):
                x = a(env)
                y = b(env)
                if x.__class__ is float and y.__class__ is float:
                    return x == y
                return leq(x, y)
            return ev_eq
        if op == '!=':
            leq = self.loose_eq
            return lambda env: not leq(a(env), b(env))
        if op == 'instanceof':
            return lambda env: vm.instanceof(a(env), b(env))
        if op == 'in':
            def ev_in(env):
                key = a(env)
                return vm.has(b(env), key)
            return ev_in
        if op == '&':
            return lambda env: float(to_int32(tonum(a(env))) & to_int32(tonum(b(env))))
        if op == '|':
            return lambda env: float(to_int32(tonum(a(env))) | to_int32(tonum(b(env))))
        if op == '^':
            return lambda env: float(to_int32(tonum(a(env))) ^ to_int32(tonum(b(env))))
        if op == '<<':
            return lambda env: float(to_int32(float(to_int32(tonum(a(env))) << (to_uint32(tonum(b(env))) & 31))))
        if op == '>>':
            return lambda env: float(to_int32(tonum(a(env))) >> (to_uint32(tonum(b(env))) & 31))
        if op == '>>>':
            return lambda env: float(to_uint32(tonum(a(env))) >> (to_uint32(tonum(b(env))) & 31))
        raise NotImplementedError('operator ' + op)

    def binop_fn(self, op):
        """function (x, y) -> value for a compound assignment operator"""
        tonum = self.tonum
        add = self.add
        table = {
            '+': lambda x, y: x + y if (x.__class__ is float and y.__class__ is float) else add(x, y),
            '-': lambda x, y: tonum(x) - tonum(y),
            '*': lambda x, y: tonum(x) * tonum(y),
            '/': lambda x, y: js_div(tonum(x), tonum(y)),
            '%': lambda x, y: js_mod(tonum(x), tonum(y)),
            '**': lambda x, y: js_pow(tonum(x), tonum(y)),
            '&': lambda x, y: float(to_int32(tonum(x)) & to_int32(tonum(y))),
            '|': lambda x, y: float(to_int32(tonum(x)) | to_int32(tonum(y))),
            '^': lambda x, y: float(to_int32(tonum(x)) ^ to_int32(tonum(y))),
            '<<': lambda x, y: float(to_int32(float(to_int32(tonum(x)) << (to_uint32(tonum(y)) & 31)))),
            '>>': lambda x, y: float(to_int32(tonum(x)) >> (to_uint32(tonum(y)) & 31)),
            '>>>': lambda x, y: float(to_uint32(tonum(x)) >> (to_uint32(tonum(y)) & 31)),
        }
        return table[op]

    def e_assign(self, node, fctx):
        _, op, target, value = node
        if op == '=':
            val = self.c_expr(value, fctx, name_hint=target[1] if target[0] == 'id' else None)
            k = target[0]
            if k == 'member':
                obj = self.c_expr(target[1], fctx)
                name = target[2]
                put = self.put

                def ev_member(env):
                    o = obj(env)
                    v = val(env)
                    if o.__class__ is JSObject:
                        o.props[name] = v
                    else:
                        put(o, name, v)
                    return v
                return ev_member
            if k == 'index':
                obj = self.c_expr(target[1], fctx)
                idx = self.c_expr(target[2], fctx)
                put = self.put

                def ev_index(env):
                    o = obj(env)
                    i = idx(env)
                    v = val(env)
                    put(o, i, v)
                    return v
                return ev_index
            bind = self.c_assign_target(target, fctx)

            def ev(env):
                v = val(env)
                bind(env, v)
                return v
            return ev
        val = self.c_expr(value, fctx)
        logical = op in ('&&=', '||=', '??=')
        fn = None if logical else self.binop_fn(op[:-1])
        k = target[0]
        get, put = self.get, self.put

        def decide(cur):
            if op == '&&=':
                return truthy(cur)
            if op == '||=':
                return not truthy(cur)
            return cur is None or cur is UNDEF
        if k == 'id':
            read = self.e_id(target, fctx)
            write = self.c_assign_target(target, fctx)

            def ev_id(env):
                cur = read(env)
                if logical:
                    if not decide(cur):
                        return cur
                    v = val(env)
                else:
                    v = fn(cur, val(env))
                write(env, v)
                return v
            return ev_id
        obj = self.c_expr(target[1], fctx)
        idx = self.c_expr(target[2], fctx) if k == 'index' else None
        name = target[2] if k == 'member' else None

        def ev_prop(env):
            o = obj(env)
            key = idx(env) if idx is not None else name
            cur = get(o, key)
            if logical:
                if not decide(cur):
                    return cur
                v = val(env)
            else:
                v = fn(cur, val(env))
            put(o, key, v)
            return v
        return ev_prop

    def e_update(self, node, fctx):
        _, op, prefix, target = node
        d = 1.0 if op == '++' else -1.0
        tonum = self.tonum
        k = target[0]
        if k == 'paren':
            return self.e_update((node[0], op, prefix, target[1]), fctx)
        if k == 'id':
            name = target[1]
            vm = self

            def ev_id(env):
                s = env
                while s is not None:
                    v = s.vars
                    if name in v:
                        old = v[name]
                        if old.__class__ is not float:
                            old = tonum(old)
                        new = old + d
                        v[name] = new
                        return new if prefix else old
                    s = s.parent
                vm.throw('ReferenceError', name + ' is not defined')
            return ev_id
        obj = self.c_expr(target[1], fctx)
        idx = self.c_expr(target[2], fctx) if k == 'index' else None
        name = target[2] if k == 'member' else None
        get, put = self.get, self.put

        def ev_prop(env):
            o = obj(env)
            key = idx(env) if idx is not None else name
            old = tonum(get(o, key))
            new = old + d
            put(o, key, new)
            return new if prefix else old
        return ev_prop

    def e_member(self, node, fctx):
        _, objn, name, optional = node
        if objn[0] == 'super':
            c = fctx.nearest_non_arrow()
            if c is not None:
                c.uses_super = True
            vm = self

            def ev_super(env):
                fn = _lookup(env, '%fn')
                return vm.get(fn.home.proto, name)
            return ev_super
        obj = self.c_expr(objn, fctx)
        get = self.get
        if optional:
            def ev_opt(env):
                o = obj(env)
                if o is None or o is UNDEF:
                    return UNDEF
                return get(o, name)
            return ev_opt

        def ev(env):
            o = obj(env)
            if o.__class__ is JSObject:
                # plain object: inline the prototype walk
                while o is not None:
                    v = o.props.get(name, _MISSING)
                    if v is not _MISSING:
                        return v
                    o = o.proto
                return UNDEF
            return get(o, name)
        return ev

    def e_index(self, node, fctx):
        _, objn, idxn, optional = node
        obj = self.c_expr(objn, fctx)
        idx = self.c_expr(idxn, fctx)
        get = self.get

        def ev(env):
            o = obj(env)
            if optional and (o is None or o is UNDEF):
                return UNDEF
            k = idx(env)
            c = o.__class__
            if (c is JSArray or c is JSTypedArray) and k.__class__ is float:
                items = o.items
                try:
                    i = int(k)
                except (ValueError, OverflowError):
                    return UNDEF
                if i == k and 0 <= i < len(items):
                    v = items[i]
                    return float(v) if v.__class__ is int else v
            return get(o, k)
        return ev

    def c_args(self, args, fctx):
        """returns ev(env) -> list of argument values"""
        if any(a[0] == 'spread' for a in args):
            parts = [(True, self.c_expr(a[1], fctx)) if a[0] == 'spread' else (False, self.c_expr(a, fctx)) for a in args]
            iterate = self.iterate

            def ev_spread(env):
                out = []
                for sp, e in parts:
                    if sp:
                        out.extend(iterate(e(env)))
                    else:
                        out.append(e(env))
                return out
            return ev_spread
        evs = tuple(self.c_expr(a, fctx) for a in args)
        n = len(evs)
        if n == 0:
            return lambda env: []
        if n == 1:
            e0 = evs[0]
            return lambda env: [e0(env)]
        if n == 2:
            e0, e1 = evs
            return lambda env: [e0(env), e1(env)]
        return lambda env: [e(env) for e in evs]

    def e_call(self, node, fctx):
        _, callee, args, optional = node
        args_ev = self.c_args(args, fctx)
        vm = self
        call = self.call
        call_js = self.call_js
        get = self.get
        k = callee[0]
        if k == 'super':
            c = fctx.nearest_non_arrow()
            if c is not None:
                c.uses_super = True

            def ev_super_call(env):
                s = env
                while s is not None:
                    if '%nt' in s.vars:
                        break
                    s = s.parent
                if s is None:
                    vm.throw('SyntaxError', "'super' keyword unexpected here")
                fn = s.vars['%fn']
                cls = fn.home.props['constructor']
                a = args_ev(env)
                this = vm.construct(cls.parent, a, s.vars['%nt'])
                s.vars['this'] = this
                if cls.fields:
                    vm.init_fields(cls, this)
                return this
            return ev_super_call
        if k == 'member' and callee[1][0] == 'super':
            c = fctx.nearest_non_arrow()
            if c is not None:
                c.uses_super = True
            name = callee[2]
            this_ev = self.e_this(('this',), fctx)

            def ev_super_method(env):
                fn = _lookup(env, '%fn')
                f = get(fn.home.proto, name)
                return call(f, this_ev(env), args_ev(env))
            return ev_super_method
        if k in ('member', 'index'):
            obj = self.c_expr(callee[1], fctx)
            name = callee[2] if k == 'member' else None
            idx = self.c_expr(callee[2], fctx) if k == 'index' else None
            opt_member = callee[3]

            def ev_method(env):
                o = obj(env)
                if opt_member and (o is None or o is UNDEF):
                    return UNDEF
                key = name if idx is None else idx(env)
                f = get(o, key)
                if optional and (f is None or f is UNDEF):
                    return UNDEF
                a = args_ev(env)
                c = f.__class__
                if c is JSFunction:
                    return call_js(f, o, a, None)
                if c is NativeFunction:
                    return f.fn(o, a)
                if f is UNDEF or f is None:
                    vm.throw('TypeError', '%s is not a function' % (key if isinstance(key, str) else vm.tostr(key)))
                return call(f, o, a)
            return ev_method
        if k == 'id' and callee[1] == 'import':
            def ev_import(env):
                a = args_ev(env)
                return vm.host_import(a[0] if a else UNDEF)
            return ev_import
        fn_ev = self.c_expr(callee, fctx)

        def ev_call(env):
            f = fn_ev(env)
            if optional and (f is None or f is UNDEF):
                return UNDEF
            a = args_ev(env)
            if f.__class__ is JSFunction:
                return call_js(f, UNDEF, a, None)
            if f is UNDEF or f is None:
                vm.throw('TypeError', '%s is not a function' % (callee[1] if k == 'id' else 'expression'))
            return call(f, UNDEF, a)
        return ev_call

    def e_new(self, node, fctx):
        _, callee, args = node
        fn_ev = self.c_expr(callee, fctx)
        args_ev = self.c_args(args, fctx)
        construct = self.construct

        def ev(env):
            f = fn_ev(env)
            return construct(f, args_ev(env), None)
        return ev

    def e_super(self, node, fctx):
        raise NotImplementedError('bare super')

    def e_spread(self, node, fctx):
        raise NotImplementedError('spread outside call / array / object')

    def e_yield(self, node, fctx):
        raise NotImplementedError('yield is supported as a statement only')

    def host_import(self, spec):
        self.throw('Error', 'dynamic import is not available: ' + self.tostr(spec))


def _array_iter(arr):
    """arrays are iterated live, by index (an array that grows while iterated is visited to its new end)"""
    i = 0
    items = arr.items
    while i < len(items):
        yield items[i]
        i += 1


def _lookup(env, name):
    s = env
    while s is not None:
        if name in s.vars:
            return s.vars[name]
        s = s.parent
    return UNDEF


def _strict_eq(a, b):
    ca = a.__class__
    if ca is float:
        return b.__class__ is float and a == b
    if ca is str:
        return b.__class__ is str and a == b
    return a is b


def _contains_function(node):
    if not isinstance(node, (tuple, list)):
        return False
    if isinstance(node, tuple) and node and node[0] in ('function', 'class'):
        return True
    for x in node:
        if isinstance(x, (tuple, list)) and _contains_function(x):
            return True
    return False


def js_div(x, y):
    try:
        return x / y
    except ZeroDivisionError:
        if x != x or x == 0:
            return math.nan
        return math.inf if (x > 0) == (math.copysign(1.0, y) > 0) else -math.inf


def js_mod(x, y):
    if y == 0 or x != x or y != y or x in (math.inf, -math.inf):
        return math.nan
    if y in (math.inf, -math.inf):
        return x
    return math.fmod(x, y)


def js_pow(x, y):
    if y != y:
        return math.nan
    if y == 0:
        return 1.0
    if (x == 1 or x == -1) and y in (math.inf, -math.inf):
        return math.nan
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf if (x > 0 or y % 2 == 0) else -math.inf
    except ValueError:
        if x == 0:
            # 0 ** negative
            if math.copysign(1.0, x) < 0 and y == int(y) and int(y) % 2 == 1:
                return -math.inf
            return math.inf
        return math.nan
    except ZeroDivisionError:
        if math.copysign(1.0, x) < 0 and y == int(y) and int(y) % 2 == 1:
            return -math.inf
        return math.inf
