"""TEST INFRASTRUCTURE — tokenizer + parser of the JavaScript subset the reference's sources use.

Part of `oracle/jsvm`: a small JavaScript interpreter whose only job is to EXECUTE THE UNMODIFIED REFERENCE
SOURCES (`/root/reference/src/*.js`, `/root/reference/tests/*/test.mjs`) in a container that has no JS engine,
so that the C++ restatement (`oracle/oracle.cpp`) can be pinned against outputs of the reference itself.
Nothing in the product path (`jsraytracer_b200/`, the GPU arm of `bench.py`) may import this package.

Grammar covered (ES2020 minus async, getters / setters, labels and tagged templates): var / let / const with
array / object destructuring and defaults, functions, generators (`function*`, `*method`), arrows, classes
(extends, static methods, static and instance fields, super calls and super property calls), template
literals, regular-expression literals, spread / rest, optional chaining, `??`, `**`, for / for-of / for-in /
while / do-while / switch / try / throw, automatic semicolon insertion.

AST: tuples `(kind, ...)`; see `Parser` methods for the shapes.
"""
import re

KEYWORDS = {
    'var', 'let', 'const', 'function', 'return', 'if', 'else', 'for', 'while', 'do', 'break', 'continue', 'new',
    'delete', 'typeof', 'instanceof', 'in', 'of', 'class', 'extends', 'super', 'this', 'null', 'true', 'false',
    'throw', 'try', 'catch', 'finally', 'switch', 'case', 'default', 'void', 'yield', 'static', 'export', 'import',
}
# words that are keywords only in some positions and otherwise identifiers
SOFT = {'of', 'static', 'let', 'yield', 'import'}

PUNCT = [
    '>>>=', '...', '===', '!==', '**=', '<<=', '>>=', '>>>', '&&=', '||=', '??=',
    '=>', '==', '!=', '<=', '>=', '&&', '||', '??', '?.', '++', '--', '+=', '-=', '*=', '/=', '%=', '&=', '|=', '^=',
    '<<', '>>', '**',
    '{', '}', '(', ')', '[', ']', ';', ',', '<', '>', '+', '-', '*', '/', '%', '&', '|', '^', '!', '~', '?', ':', '=', '.',
]

_num_re = re.compile(r'0[xX][0-9a-fA-F]+|0[bB][01]+|0[oO][0-7]+|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)')
_id_re = re.compile(r'[A-Za-z_$][A-Za-z0-9_$]*')
_ws_re = re.compile(r'[ \t\r\f\v﻿]+')


class JSSyntaxError(Exception):
    pass


class Tok:
    __slots__ = ('t', 'v', 'nl', 'pos')

    def __init__(self, t, v, nl, pos):
        self.t = t      # 'num' 'str' 'id' 'kw' 'p' 'tmpl' 're' 'eof'
        self.v = v
        self.nl = nl    # a line terminator precedes this token
        self.pos = pos

    def __repr__(self):
        return '%s:%r' % (self.t, self.v)


_ESC = {'n': '\n', 't': '\t', 'r': '\r', 'b': '\b', 'f': '\f', 'v': '\v', '0': '\0'}


def _read_string(src, i, quote):
    out = []
    n = len(src)
    while i < n:
        c = src[i]
        if c == quote:
            return ''.join(out), i + 1
        if c == '\\':
            i += 1
            c = src[i]
            if c == 'x':
                out.append(chr(int(src[i + 1:i + 3], 16))); i += 3; continue
            if c == 'u':
                if src[i + 1] == '{':
                    j = src.index('}', i)
                    out.append(chr(int(src[i + 2:j], 16))); i = j + 1; continue
                out.append(chr(int(src[i + 1:i + 5], 16))); i += 5; continue
            if c == '\n':
                i += 1; continue
            out.append(_ESC.get(c, c)); i += 1; continue
        out.append(c); i += 1
    raise JSSyntaxError('unterminated string')


def tokenize(src):
    toks = []
    i, n = 0, len(src)
    nl = False
    while i < n:
        c = src[i]
        if c == '\n':
            nl = True; i += 1; continue
        m = _ws_re.match(src, i)
        if m:
            i = m.end(); continue
        if c == '/' and i + 1 < n:
            d = src[i + 1]
            if d == '/':
                j = src.find('\n', i)
                i = n if j < 0 else j
                continue
            if d == '*':
                j = src.find('*/', i + 2)
                if j < 0:
                    raise JSSyntaxError('unterminated comment')
                if '\n' in src[i:j]:
                    nl = True
                i = j + 2
                continue
        if c.isdigit() or (c == '.' and i + 1 < n and src[i + 1].isdigit()):
            m = _num_re.match(src, i)
            s = m.group(0)
            if s[:2] in ('0x', '0X'):
                v = float(int(s, 16))
            elif s[:2] in ('0b', '0B'):
                v = float(int(s[2:], 2))
            elif s[:2] in ('0o', '0O'):
                v = float(int(s[2:], 8))
            else:
                v = float(s)
            toks.append(Tok('num', v, nl, i)); nl = False; i = m.end(); continue
        m = _id_re.match(src, i)
        if m:
            s = m.group(0)
            toks.append(Tok('kw' if s in KEYWORDS else 'id', s, nl, i)); nl = False; i = m.end(); continue
        if c == '"' or c == "'":
            s, j = _read_string(src, i + 1, c)
            toks.append(Tok('str', s, nl, i)); nl = False; i = j; continue
        if c == '`':
            # template literal: (cooked strings, expression sources)
            strs, exprs, cur = [], [], []
            j = i + 1
            while True:
                if j >= n:
                    raise JSSyntaxError('unterminated template')
                ch = src[j]
                if ch == '`':
                    j += 1; break
                if ch == '\\':
                    nxt = src[j + 1]
                    cur.append(_ESC.get(nxt, nxt)); j += 2; continue
                if ch == '$' and j + 1 < n and src[j + 1] == '{':
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == '{': depth += 1
                        elif src[k] == '}': depth -= 1
                        k += 1
                    strs.append(''.join(cur)); cur = []
                    exprs.append(src[j + 2:k - 1]); j = k; continue
                cur.append(ch); j += 1
            strs.append(''.join(cur))
            toks.append(Tok('tmpl', (strs, exprs), nl, i)); nl = False; i = j; continue
        if c == '/':
            # regular expression or division: decided by the previous token
            prev = toks[-1] if toks else None
            is_div = prev is not None and (prev.t in ('num', 'str', 'id', 'tmpl', 're') or
                                           (prev.t == 'p' and prev.v in (')', ']', '}')) or
                                           (prev.t == 'kw' and prev.v in ('this', 'super', 'null', 'true', 'false')))
            if not is_div:
                j, in_class = i + 1, False
                while True:
                    ch = src[j]
                    if ch == '\\': j += 2; continue
                    if ch == '[': in_class = True
                    elif ch == ']': in_class = False
                    elif ch == '/' and not in_class: break
                    elif ch == '\n': raise JSSyntaxError('unterminated regex')
                    j += 1
                body = src[i + 1:j]
                m = _id_re.match(src, j + 1)
                flags = m.group(0) if m else ''
                toks.append(Tok('re', (body, flags), nl, i)); nl = False
                i = j + 1 + len(flags); continue
        for p in PUNCT:
            if src.startswith(p, i):
                toks.append(Tok('p', p, nl, i)); nl = False; i += len(p); break
        else:
            raise JSSyntaxError('unexpected character %r at %d' % (c, i))
    toks.append(Tok('eof', None, nl, n))
    return toks


BINPREC = {
    '??': 4, '||': 5, '&&': 6, '|': 7, '^': 8, '&': 9,
    '==': 10, '!=': 10, '===': 10, '!==': 10,
    '<': 11, '>': 11, '<=': 11, '>=': 11, 'instanceof': 11, 'in': 11,
    '<<': 12, '>>': 12, '>>>': 12, '+': 13, '-': 13, '*': 14, '/': 14, '%': 14, '**': 15,
}
ASSIGN_OPS = {'=', '+=', '-=', '*=', '/=', '%=', '**=', '<<=', '>>=', '>>>=', '&=', '|=', '^=', '&&=', '||=', '??='}


class Parser:
    def __init__(self, src, filename='<js>'):
        self.src = src
        self.filename = filename
        self.toks = tokenize(src)
        self.i = 0
        self.no_in = False

    # -- token helpers ------------------------------------------------------------------------
    def err(self, msg):
        t = self.toks[self.i]
        line = self.src.count('\n', 0, t.pos) + 1
        raise JSSyntaxError('%s:%d: %s (at %r)' % (self.filename, line, msg, t.v))

    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def is_p(self, v, k=0):
        t = self.peek(k)
        return t.t == 'p' and t.v == v

    def is_kw(self, v, k=0):
        t = self.peek(k)
        return t.t == 'kw' and t.v == v

    def eat_p(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def eat_kw(self, v):
        if self.is_kw(v):
            self.i += 1
            return True
        return False

    def expect_p(self, v):
        if not self.eat_p(v):
            self.err('expected %r' % v)

    def ident(self):
        t = self.next()
        if t.t == 'id' or (t.t == 'kw' and t.v in SOFT):
            return t.v
        self.i -= 1
        self.err('expected identifier')

    def prop_name(self):
        """identifier-name (keywords allowed), string or number: returns a str"""
        t = self.next()
        if t.t in ('id', 'kw', 'str'):
            return t.v
        if t.t == 'num':
            from .runtime import num_to_str
            return num_to_str(t.v)
        self.i -= 1
        self.err('expected property name')

    def semicolon(self):
        if self.eat_p(';'):
            return
        t = self.peek()
        if t.t == 'eof' or t.nl or (t.t == 'p' and t.v == '}'):
            return
        self.err('expected ;')

    # -- program / statements -----------------------------------------------------------------
    def parse_program(self):
        body = []
        while self.peek().t != 'eof':
            body.append(self.statement())
        return ('program', body)

    def statement(self):
        t = self.peek()
        if t.t == 'p':
            if t.v == '{':
                return self.block()
            if t.v == ';':
                self.next()
                return ('empty',)
        elif t.t == 'kw':
            v = t.v
            if v in ('var', 'const') or (v == 'let' and (self.peek(1).t in ('id',) or self.is_p('[', 1) or self.is_p('{', 1))):
                d = self.var_decl()
                self.semicolon()
                return d
            if v == 'function':
                return self.function(True)
            if v == 'class':
                return self.klass(True)
            if v == 'return':
                self.next()
                nt = self.peek()
                arg = None
                if not (nt.nl or nt.t == 'eof' or (nt.t == 'p' and nt.v in (';', '}'))):
                    arg = self.expression()
                self.semicolon()
                return ('return', arg)
            if v == 'if':
                self.next(); self.expect_p('(')
                test = self.expression()
                self.expect_p(')')
                cons = self.statement()
                alt = self.statement() if self.eat_kw('else') else None
                return ('if', test, cons, alt)
            if v == 'for':
                return self.for_stmt()
            if v == 'while':
                self.next(); self.expect_p('(')
                test = self.expression()
                self.expect_p(')')
                return ('while', test, self.statement())
            if v == 'do':
                self.next()
                body = self.statement()
                if not self.eat_kw('while'):
                    self.err('expected while')
                self.expect_p('(')
                test = self.expression()
                self.expect_p(')')
                self.eat_p(';')
                return ('dowhile', body, test)
            if v in ('break', 'continue'):
                self.next()
                nt = self.peek()
                if nt.t == 'id' and not nt.nl:
                    self.err('labels are not supported')
                self.semicolon()
                return (v,)
            if v == 'throw':
                self.next()
                arg = self.expression()
                self.semicolon()
                return ('throw', arg)
            if v == 'try':
                self.next()
                blk = self.block()
                param = handler = final = None
                if self.eat_kw('catch'):
                    if self.eat_p('('):
                        param = self.binding_target()
                        self.expect_p(')')
                    handler = self.block()
                if self.eat_kw('finally'):
                    final = self.block()
                return ('try', blk, param, handler, final)
            if v == 'switch':
                return self.switch_stmt()
            if v == 'export':
                self.next()
                self.eat_kw('default')
                return self.statement()
            if v == 'import' and not self.is_p('(', 1):
                # static imports are resolved by the host (sources are loaded as classic scripts): skip the statement
                while not self.is_p(';') and not self.peek().nl:
                    self.next()
                    if self.peek().t == 'eof':
                        break
                self.eat_p(';')
                return ('empty',)
        e = self.expression()
        self.semicolon()
        return ('expr', e)

    def block(self):
        self.expect_p('{')
        body = []
        while not self.is_p('}'):
            if self.peek().t == 'eof':
                self.err('unterminated block')
            body.append(self.statement())
        self.next()
        return ('block', body)

    def var_decl(self):
        kind = self.next().v
        decls = []
        while True:
            target = self.binding_target()
            init = None
            if self.eat_p('='):
                init = self.assignment()
            decls.append((target, init))
            if not self.eat_p(','):
                break
        return ('vardecl', kind, decls)

    def binding_target(self):
        if self.is_p('[') or self.is_p('{'):
            return self.to_pattern(self.primary())
        return ('id', self.ident())

    def for_stmt(self):
        self.next()
        self.expect_p('(')
        init = None
        if self.is_p(';'):
            pass
        elif self.is_kw('var') or self.is_kw('const') or self.is_kw('let'):
            kind = self.peek().v
            save = self.i
            self.next()
            target = self.binding_target()
            if self.eat_kw('of'):
                it = self.assignment()
                self.expect_p(')')
                return ('forof', kind, target, it, self.statement())
            if self.eat_kw('in'):
                obj = self.expression()
                self.expect_p(')')
                return ('forin', kind, target, obj, self.statement())
            self.i = save
            self.no_in = True
            init = self.var_decl()
            self.no_in = False
        else:
            save = self.i
            self.no_in = True
            e = self.expression()
            self.no_in = False
            if self.eat_kw('of'):
                it = self.assignment()
                self.expect_p(')')
                return ('forof', None, self.to_pattern(e), it, self.statement())
            if self.eat_kw('in'):
                obj = self.expression()
                self.expect_p(')')
                return ('forin', None, self.to_pattern(e), obj, self.statement())
            init = ('expr', e)
        self.expect_p(';')
        test = None if self.is_p(';') else self.expression()
        self.expect_p(';')
        update = None if self.is_p(')') else self.expression()
        self.expect_p(')')
        return ('for', init, test, update, self.statement())

    def switch_stmt(self):
        self.next()
        self.expect_p('(')
        disc = self.expression()
        self.expect_p(')')
        self.expect_p('{')
        cases = []
        while not self.eat_p('}'):
            if self.eat_kw('default'):
                test = None
            else:
                if not self.eat_kw('case'):
                    self.err('expected case')
                test = self.expression()
            self.expect_p(':')
            body = []
            while not (self.is_kw('case') or self.is_kw('default') or self.is_p('}')):
                body.append(self.statement())
            cases.append((test, body))
        return ('switch', disc, cases)

    # -- functions and classes ----------------------------------------------------------------
    def params(self):
        """after '(' ... returns list of patterns (last may be ('rest', pattern)) and consumes ')'"""
        ps = []
        while not self.is_p(')'):
            if self.eat_p('...'):
                ps.append(('rest', self.binding_target()))
            else:
                target = self.binding_target()
                if self.eat_p('='):
                    target = ('defpat', target, self.assignment())
                ps.append(target)
            if not self.eat_p(','):
                break
        self.expect_p(')')
        return ps

    def function(self, is_decl):
        self.next()  # 'function'
        gen = self.eat_p('*')
        name = None
        if not self.is_p('('):
            name = self.ident()
        self.expect_p('(')
        ps = self.params()
        body = self.block()
        node = ('function', name, ps, body, {'gen': gen, 'arrow': False, 'expr_body': False})
        return ('funcdecl', name, node) if is_decl else node

    def method_function(self, name, gen=False):
        self.expect_p('(')
        ps = self.params()
        body = self.block()
        return ('function', name, ps, body, {'gen': gen, 'arrow': False, 'expr_body': False, 'method': True})

    def klass(self, is_decl):
        self.next()  # 'class'
        name = None
        if not self.is_kw('extends') and not self.is_p('{'):
            name = self.ident()
        parent = None
        if self.eat_kw('extends'):
            parent = self.unary_postfix_member()
        self.expect_p('{')
        ctor = None
        methods = []   # (static, key_expr_or_str, computed, function)
        fields = []    # (static, key, computed, init_expr or None)
        while not self.eat_p('}'):
            if self.eat_p(';'):
                continue
            static = False
            if self.is_kw('static') and not (self.is_p('(', 1) or self.is_p('=', 1)):
                self.next()
                static = True
            gen = self.eat_p('*')
            if (self.peek().t == 'id' and self.peek().v in ('get', 'set') and
                    not (self.is_p('(', 1) or self.is_p('=', 1) or self.is_p(';', 1) or self.is_p('}', 1))):
                self.err('getters / setters are not supported')
            computed = False
            if self.eat_p('['):
                key = self.assignment()
                self.expect_p(']')
                computed = True
            else:
                key = self.prop_name()
            if self.is_p('('):
                fn = self.method_function(None if computed else key, gen)
                if key == 'constructor' and not static and not computed:
                    ctor = fn
                else:
                    methods.append((static, key, computed, fn))
            else:
                init = None
                if self.eat_p('='):
                    init = self.assignment()
                self.semicolon()
                fields.append((static, key, computed, init))
        node = ('class', name, parent, ctor, methods, fields)
        return ('classdecl', name, node) if is_decl else node

    # -- expressions --------------------------------------------------------------------------
    def expression(self):
        e = self.assignment()
        if self.is_p(','):
            es = [e]
            while self.eat_p(','):
                es.append(self.assignment())
            return ('seq', es)
        return e

    def is_arrow_ahead(self):
        """at '(' : is this the parameter list of an arrow function?"""
        depth, k = 0, 0
        while True:
            t = self.peek(k)
            if t.t == 'eof':
                return False
            if t.t == 'p':
                if t.v in ('(', '[', '{'):
                    depth += 1
                elif t.v in (')', ']', '}'):
                    depth -= 1
                    if depth == 0:
                        nt = self.peek(k + 1)
                        return nt.t == 'p' and nt.v == '=>'
            k += 1

    def arrow_body(self, ps):
        if self.is_p('{'):
            body = self.block()
            return ('function', None, ps, body, {'gen': False, 'arrow': True, 'expr_body': False})
        save = self.no_in
        self.no_in = False
        body = self.assignment()
        self.no_in = save
        return ('function', None, ps, body, {'gen': False, 'arrow': True, 'expr_body': True})

    def assignment(self):
        t = self.peek()
        # arrow functions
        if t.t == 'id' and self.is_p('=>', 1):
            name = self.next().v
            self.next()
            return self.arrow_body([('id', name)])
        if t.t == 'p' and t.v == '(' and self.is_arrow_ahead():
            self.next()
            ps = self.params()
            self.expect_p('=>')
            return self.arrow_body(ps)
        if t.t == 'kw' and t.v == 'yield':
            self.next()
            nt = self.peek()
            delegate = False
            arg = None
            if not (nt.nl or nt.t == 'eof' or (nt.t == 'p' and nt.v in (';', '}', ')', ']', ','))):
                delegate = self.eat_p('*')
                arg = self.assignment()
            return ('yield', arg, delegate)
        left = self.conditional()
        t = self.peek()
        if t.t == 'p' and t.v in ASSIGN_OPS:
            self.next()
            right = self.assignment()
            if t.v == '=':
                return ('assign', '=', self.to_pattern(left), right)
            if left[0] not in ('id', 'member', 'index'):
                self.err('invalid assignment target')
            return ('assign', t.v, left, right)
        return left

    def conditional(self):
        test = self.binary(0)
        if self.eat_p('?'):
            save = self.no_in
            self.no_in = False
            cons = self.assignment()
            self.no_in = save
            self.expect_p(':')
            alt = self.assignment()
            return ('cond', test, cons, alt)
        return test

    def binary(self, minprec):
        left = self.unary()
        while True:
            t = self.peek()
            if t.t == 'p' or (t.t == 'kw' and t.v in ('instanceof', 'in')):
                op = t.v
                if op == 'in' and self.no_in:
                    break
                prec = BINPREC.get(op)
                if prec is None or prec <= minprec:
                    break
                self.next()
                right = self.binary(prec - 1 if op == '**' else prec)
                if op in ('&&', '||', '??'):
                    left = ('logical', op, left, right)
                else:
                    left = ('bin', op, left, right)
            else:
                break
        return left

    def unary(self):
        t = self.peek()
        if t.t == 'p':
            if t.v in ('!', '-', '+', '~'):
                self.next()
                return ('unary', t.v, self.unary())
            if t.v in ('++', '--'):
                self.next()
                return ('update', t.v, True, self.unary())
        elif t.t == 'kw' and t.v in ('typeof', 'void', 'delete'):
            self.next()
            return ('unary', t.v, self.unary())
        return self.postfix()

    def postfix(self):
        e = self.unary_postfix_member()
        t = self.peek()
        if t.t == 'p' and t.v in ('++', '--') and not t.nl:
            self.next()
            return ('update', t.v, False, e)
        return e

    def arguments(self):
        args = []
        while not self.is_p(')'):
            if self.eat_p('...'):
                args.append(('spread', self.assignment()))
            else:
                args.append(self.assignment())
            if not self.eat_p(','):
                break
        self.expect_p(')')
        return args

    def unary_postfix_member(self):
        """member / call / new chain"""
        if self.is_kw('new'):
            self.next()
            if self.is_p('.'):
                self.err('new.target is not supported')
            callee = self.member_only()
            args = self.arguments() if self.eat_p('(') else []
            e = ('new', callee, args)
        else:
            e = self.primary()
        while True:
            t = self.peek()
            if t.t == 'p':
                if t.v == '.':
                    self.next()
                    e = ('member', e, self.prop_name(), False)
                    continue
                if t.v == '?.':
                    self.next()
                    if self.eat_p('('):
                        e = ('call', e, self.arguments(), True)
                    elif self.eat_p('['):
                        idx = self.expression()
                        self.expect_p(']')
                        e = ('index', e, idx, True)
                    else:
                        e = ('member', e, self.prop_name(), True)
                    continue
                if t.v == '[':
                    self.next()
                    save = self.no_in
                    self.no_in = False
                    idx = self.expression()
                    self.no_in = save
                    self.expect_p(']')
                    e = ('index', e, idx, False)
                    continue
                if t.v == '(':
                    self.next()
                    save = self.no_in
                    self.no_in = False
                    args = self.arguments()
                    self.no_in = save
                    e = ('call', e, args, False)
                    continue
            elif t.t == 'tmpl':
                self.err('tagged templates are not supported')
            break
        return e

    def member_only(self):
        """callee of `new`: member accesses but no calls"""
        if self.is_kw('new'):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.eat_p('(') else []
            e = ('new', callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat_p('.'):
                e = ('member', e, self.prop_name(), False)
            elif self.is_p('['):
                self.next()
                idx = self.expression()
                self.expect_p(']')
                e = ('index', e, idx, False)
            else:
                break
        return e

    def primary(self):
        t = self.next()
        tt = t.t
        if tt == 'num':
            return ('num', t.v)
        if tt == 'str':
            return ('str', t.v)
        if tt == 'id':
            return ('id', t.v)
        if tt == 'tmpl':
            strs, exprs = t.v
            return ('template', strs, [Parser(s, self.filename).parse_expression_only() for s in exprs])
        if tt == 're':
            return ('regex', t.v[0], t.v[1])
        if tt == 'p':
            v = t.v
            if v == '(':
                save = self.no_in
                self.no_in = False
                e = self.expression()
                self.no_in = save
                self.expect_p(')')
                return ('paren', e)
            if v == '[':
                save = self.no_in
                self.no_in = False
                elems = []
                while not self.is_p(']'):
                    if self.is_p(','):
                        self.next()
                        elems.append(None)
                        continue
                    if self.eat_p('...'):
                        elems.append(('spread', self.assignment()))
                    else:
                        elems.append(self.assignment())
                    if not self.is_p(']'):
                        self.expect_p(',')
                self.next()
                self.no_in = save
                return ('array', elems)
            if v == '{':
                save = self.no_in
                self.no_in = False
                r = self.object_literal()
                self.no_in = save
                return r
        elif tt == 'kw':
            v = t.v
            if v == 'this':
                return ('this',)
            if v == 'null':
                return ('null',)
            if v == 'true':
                return ('bool', True)
            if v == 'false':
                return ('bool', False)
            if v == 'function':
                self.i -= 1
                return self.function(False)
            if v == 'class':
                self.i -= 1
                return self.klass(False)
            if v == 'super':
                return ('super',)
            if v == 'import' and self.is_p('('):
                return ('id', 'import')
            if v in SOFT:
                return ('id', v)
        self.i -= 1
        self.err('unexpected token')

    def parse_expression_only(self):
        e = self.expression()
        if self.peek().t != 'eof':
            self.err('trailing tokens in template expression')
        return e

    def object_literal(self):
        props = []   # ('prop', key, computed, value) | ('spread', expr) | ('shorthand_default', name, default)
        while not self.is_p('}'):
            if self.eat_p('...'):
                props.append(('spread', self.assignment()))
            else:
                gen = self.eat_p('*')
                computed = False
                t = self.peek()
                if self.eat_p('['):
                    key = self.assignment()
                    self.expect_p(']')
                    computed = True
                else:
                    if (t.t == 'id' and t.v in ('get', 'set') and not (self.is_p(',', 1) or self.is_p(':', 1) or
                                                                      self.is_p('(', 1) or self.is_p('}', 1) or self.is_p('=', 1))):
                        self.err('getters / setters are not supported')
                    key = self.prop_name()
                if self.is_p('('):
                    props.append(('prop', key, computed, self.method_function(None if computed else key, gen)))
                elif self.eat_p(':'):
                    props.append(('prop', key, computed, self.assignment()))
                elif self.eat_p('='):
                    props.append(('shorthand_default', key, self.assignment()))
                else:
                    if t.t not in ('id', 'kw'):
                        self.err('bad shorthand property')
                    props.append(('prop', key, False, ('id', key)))
            if not self.is_p('}'):
                self.expect_p(',')
        self.next()
        return ('object', props)

    # -- patterns -----------------------------------------------------------------------------
    def to_pattern(self, e):
        k = e[0]
        if k in ('id', 'member', 'index'):
            return e
        if k == 'paren':
            return self.to_pattern(e[1])
        if k == 'assign' and e[1] == '=':
            return ('defpat', e[2], e[3])
        if k == 'defpat' or k == 'arrpat' or k == 'objpat':
            return e
        if k == 'array':
            elems, rest = [], None
            for x in e[1]:
                if x is None:
                    elems.append(None)
                elif x[0] == 'spread':
                    rest = self.to_pattern(x[1])
                else:
                    elems.append(self.to_pattern(x))
            return ('arrpat', elems, rest)
        if k == 'object':
            props, rest = [], None
            for p in e[1]:
                if p[0] == 'spread':
                    rest = self.to_pattern(p[1])
                elif p[0] == 'shorthand_default':
                    props.append((p[1], False, ('defpat', ('id', p[1]), p[2])))
                else:
                    props.append((p[1], p[2], self.to_pattern(p[3])))
            return ('objpat', props, rest)
        self.err('invalid destructuring target')


def parse(src, filename='<js>'):
    return Parser(src, filename).parse_program()
