"""A minimal Erlang evaluator -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).

Purpose: the build image has no Erlang/OTP, so the reference cannot run here.  This module executes the reference's
UNMODIFIED source files (src/fmi.erl, bwa.erl, sga.erl, sw.erl + include/bwt.hrl, read from /root/reference at generation
time, never copied) with a small tree-walking evaluator of the Erlang subset they use: function clauses with patterns and
guards, case / if / fun / list comprehensions, tuples, cons lists, characters, strings, integer and float arithmetic,
term comparison in Erlang's term order, proper tail calls.  The handful of OTP library functions the four modules call
(lists:reverse/max/nth/nthtail/seq/sort/sublist/foldl, element, size, length, trunc, integer_to_list, io:format as a
no-op) are re-implemented here from their documented semantics.  It is NOT a BEAM: what it pins is that the hand-written
oracle (oracle/bwt_oracle.py, bwt_oracle.c) and the CUDA path agree with the reference's own source text as a generic
evaluator reads it, clause order, guards, term order of lists:max/1 and all -- not OTP's implementation of lists:sort/2
(the top-1 tie order of worker_bwt.erl:104 stays unpinned).

scripts/make_reference_golden_interpreted.py drives it; tests/golden/reference_outputs_interpreted.json is its output.
"""
from __future__ import annotations

import os
import re
import sys
from typing import Any, Dict, List, Optional, Tuple

sys.setrecursionlimit(20000)


# ---------------------------------------------------------------------------------------------------------------------
# terms
# ---------------------------------------------------------------------------------------------------------------------
class Atom(str):
    __slots__ = ()

    def __repr__(self):
        return "'%s'" % str.__str__(self)


_ATOMS: Dict[str, Atom] = {}


def atom(name: str) -> Atom:
    a = _ATOMS.get(name)
    if a is None:
        a = _ATOMS[name] = Atom(name)
    return a


class Nil:
    __slots__ = ()

    def __repr__(self):
        return "[]"


NIL = Nil()


class Cons:
    __slots__ = ("h", "t")

    def __init__(self, h, t):
        self.h = h
        self.t = t

    def __repr__(self):
        return "[%s]" % ",".join(repr(x) for x in to_pylist(self))


class Fun:
    __slots__ = ("clauses", "env", "mod")

    def __init__(self, clauses, env, mod):
        self.clauses, self.env, self.mod = clauses, env, mod


class ErlError(Exception):
    pass


class ReadDevice:
    """a file opened with [read] in list mode: file:read_line/1 returns the next line with its newline, or eof"""
    __slots__ = ("d", "p")

    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def read_line(self):
        if self.p >= len(self.d):
            return None
        k = self.d.find(b"\n", self.p)
        k = len(self.d) if k < 0 else k + 1
        ln = self.d[self.p:k]
        self.p = k
        return ln


class Device:
    """what file:open/2 returns here: io:format/3 appends to it (seeds:generate_fs/3 WRITES an Erlang module, seeds.erl:13-47)"""
    __slots__ = ("name", "buf")

    def __init__(self, name):
        self.name, self.buf = name, []

    def text(self) -> str:
        return "".join(self.buf)


def from_pylist(xs) -> Any:
    out = NIL
    for x in reversed(list(xs)):
        out = Cons(x, out)
    return out


def to_pylist(l) -> list:
    out = []
    while l is not NIL:
        if not isinstance(l, Cons):
            raise ErlError("badarg: improper list")
        out.append(l.h)
        l = l.t
    return out


def from_bytes(b: bytes):
    return from_pylist(b)


TRUE, FALSE, OK = atom("true"), atom("false"), atom("ok")


def _rank(x) -> int:
    # number < atom < reference < fun < port < pid < tuple < map < nil < list < bitstring
    if isinstance(x, bool):
        raise ErlError("python bool leaked into a term")
    if isinstance(x, (int, float)):
        return 0
    if isinstance(x, Atom):
        return 1
    if isinstance(x, (Fun, Device, ReadDevice)):
        return 3
    if isinstance(x, tuple):
        return 6
    if x is NIL:
        return 8
    if isinstance(x, Cons):
        return 9
    raise ErlError("unknown term %r" % (x,))


def compare(a, b) -> int:
    """Erlang term order: -1, 0, 1 (== semantics: 1 == 1.0)"""
    ra, rb = _rank(a), _rank(b)
    if ra != rb:
        return -1 if ra < rb else 1
    if ra == 0:
        return -1 if a < b else (1 if a > b else 0)
    if ra == 1:
        sa, sb = str.__str__(a), str.__str__(b)
        return -1 if sa < sb else (1 if sa > sb else 0)
    if ra == 6:
        if len(a) != len(b):
            return -1 if len(a) < len(b) else 1
        for x, y in zip(a, b):
            c = compare(x, y)
            if c:
                return c
        return 0
    if ra == 8:
        return 0
    if ra == 9:
        while True:
            c = compare(a.h, b.h)
            if c:
                return c
            a, b = a.t, b.t
            if a is NIL or b is NIL or not isinstance(a, Cons) or not isinstance(b, Cons):
                return compare(a, b)
    return 0


def exact_eq(a, b) -> bool:
    """=:= : like == but 1 =/= 1.0"""
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return type(a) is type(b) and a == b
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(exact_eq(x, y) for x, y in zip(a, b))
    if isinstance(a, Cons) and isinstance(b, Cons):
        while isinstance(a, Cons) and isinstance(b, Cons):
            if not exact_eq(a.h, b.h):
                return False
            a, b = a.t, b.t
        return exact_eq(a, b)
    if _rank(a) != _rank(b):
        return False
    return compare(a, b) == 0


# ---------------------------------------------------------------------------------------------------------------------
# tokens
# ---------------------------------------------------------------------------------------------------------------------
KEYWORDS = {"case", "of", "end", "if", "fun", "when", "begin", "div", "rem", "band", "bor", "bxor", "bsl", "bsr", "and", "or",
            "not", "andalso", "orelse", "xor", "bnot", "receive", "after", "try", "catch"}
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|%[^\n]*)
  | (?P<float>\d+\.\d+(?:[eE][+-]?\d+)?)
  | (?P<based>\d+\#[0-9a-zA-Z]+)
  | (?P<int>\d+)
  | (?P<char>\$(?:\\.|.))
  | (?P<string>"(?:\\.|[^"\\])*")
  | (?P<qatom>'(?:\\.|[^'\\])*')
  | (?P<var>[A-Z_][A-Za-z0-9_@]*)
  | (?P<atom>[a-z][A-Za-z0-9_@]*)
  | (?P<op>->|\+\+|--|=:=|=/=|==|/=|=<|>=|<-|\|\||<<|>>|::|[()\[\]{},;.|=<>+\-*/!?:\#])
""", re.X | re.S)
ESC = {"n": 10, "t": 9, "r": 13, "s": 32, "\\": 92, '"': 34, "'": 39, "0": 0, "e": 27}


def _unescape(s: str) -> List[int]:
    out, i = [], 0
    while i < len(s):
        if s[i] == "\\" and i + 1 < len(s):
            out.append(ESC.get(s[i + 1], ord(s[i + 1])))
            i += 2
        else:
            out.append(ord(s[i]))
            i += 1
    return out


def tokenize(text: str) -> List[Tuple[str, Any]]:
    toks, pos = [], 0
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise SyntaxError("cannot tokenize at %r" % text[pos:pos + 30])
        pos = m.end()
        k = m.lastgroup
        v = m.group(k)
        if k == "ws":
            continue
        if k == "float":
            toks.append(("float", float(v)))
        elif k == "based":
            base, digits = v.split("#")
            toks.append(("int", int(digits, int(base))))
        elif k == "int":
            toks.append(("int", int(v)))
        elif k == "char":
            body = v[1:]
            toks.append(("int", _unescape(body)[0]))
        elif k == "string":
            toks.append(("string", _unescape(v[1:-1])))
        elif k == "qatom":
            toks.append(("atom", "".join(map(chr, _unescape(v[1:-1])))))
        elif k == "var":
            toks.append(("var", v))
        elif k == "atom":
            toks.append(("kw", v) if v in KEYWORDS else ("atom", v))
        else:
            toks.append(("op", v))
    return toks


# ---------------------------------------------------------------------------------------------------------------------
# parser -> AST tuples
# ---------------------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks, macros):
        self.t, self.i, self.macros = toks, 0, macros

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        tok = self.peek()
        if tok[0] == kind and (val is None or tok[1] == val):
            self.i += 1
            return tok
        return None

    def expect(self, kind, val=None):
        tok = self.accept(kind, val)
        if tok is None:
            raise SyntaxError("expected %s %r, got %r near token %d" % (kind, val, self.peek(), self.i))
        return tok

    # -- forms ---------------------------------------------------------------------------------------------------
    def forms(self):
        funs: Dict[Tuple[str, int], list] = {}
        while self.peek()[0] != "eof":
            if self.accept("op", "-"):                       # attribute
                name = self.expect("atom")[1]
                if name == "define":
                    self.expect("op", "(")
                    mname = self.next()[1]
                    self.expect("op", ",")
                    start = self.i
                    depth = 0
                    while True:                              # the macro body: tokens up to the closing ")."
                        tok = self.peek()
                        if tok == ("op", "(") or tok == ("op", "[") or tok == ("op", "{"):
                            depth += 1
                        elif tok == ("op", ")") or tok == ("op", "]") or tok == ("op", "}"):
                            if depth == 0:
                                break
                            depth -= 1
                        self.i += 1
                    self.macros[mname] = self.t[start:self.i]
                    self.expect("op", ")")
                    self.expect("op", ".")
                else:                                        # module / export / include (handled by the loader) / others
                    while not (self.peek() == ("op", ".") and self._form_end()):
                        self.i += 1
                    self.expect("op", ".")
                continue
            name = self.expect("atom")[1]
            clauses = []
            while True:
                clauses.append(self.clause())
                if self.accept("op", ";"):
                    n2 = self.expect("atom")[1]
                    if n2 != name:
                        raise SyntaxError("clause of %s inside %s" % (n2, name))
                    continue
                self.expect("op", ".")
                break
            arity = len(clauses[0][0])
            funs.setdefault((name, arity), []).extend(clauses)
        return funs

    def _form_end(self):
        nxt = self.peek(1)
        return nxt[0] == "eof" or nxt == ("op", "-") or nxt[0] == "atom"

    def clause(self):
        self.expect("op", "(")
        pats = self.exprs_until(")")
        guard = self.guard() if self.accept("kw", "when") else None
        self.expect("op", "->")
        return (pats, guard, self.body())

    def guard(self):
        alts = [[self.expr()]]
        while True:
            if self.accept("op", ","):
                alts[-1].append(self.expr())
            elif self.accept("op", ";"):
                alts.append([self.expr()])
            else:
                return alts

    def body(self):
        es = [self.expr()]
        while self.accept("op", ","):
            es.append(self.expr())
        return es

    def exprs_until(self, close):
        out = []
        if self.accept("op", close):
            return out
        while True:
            out.append(self.expr())
            if self.accept("op", ","):
                continue
            self.expect("op", close)
            return out

    # -- expressions, by precedence --------------------------------------------------------------------------------
    def expr(self):
        left = self.e_orelse()
        if self.accept("op", "="):
            return ("match", left, self.expr())
        return left

    def e_orelse(self):
        left = self.e_andalso()
        while self.accept("kw", "orelse"):
            left = ("orelse", left, self.e_andalso())
        return left

    def e_andalso(self):
        left = self.e_cmp()
        while self.accept("kw", "andalso"):
            left = ("andalso", left, self.e_cmp())
        return left

    def e_cmp(self):
        left = self.e_listop()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("==", "/=", "=<", "<", ">=", ">", "=:=", "=/="):
            self.i += 1
            return ("op2", tok[1], left, self.e_listop())
        return left

    def e_listop(self):
        left = self.e_add()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("++", "--"):
            self.i += 1
            return ("op2", tok[1], left, self.e_listop())      # right associative
        return left

    def e_add(self):
        left = self.e_mul()
        while True:
            tok = self.peek()
            if (tok[0] == "op" and tok[1] in ("+", "-")) or (tok[0] == "kw" and tok[1] in ("bor", "bxor", "bsl", "bsr", "or", "xor")):
                self.i += 1
                left = ("op2", tok[1], left, self.e_mul())
            else:
                return left

    def e_mul(self):
        left = self.e_unary()
        while True:
            tok = self.peek()
            if (tok[0] == "op" and tok[1] in ("*", "/")) or (tok[0] == "kw" and tok[1] in ("div", "rem", "band", "and")):
                self.i += 1
                left = ("op2", tok[1], left, self.e_unary())
            else:
                return left

    def e_unary(self):
        tok = self.peek()
        if (tok[0] == "op" and tok[1] in ("-", "+")) or (tok[0] == "kw" and tok[1] in ("not", "bnot")):
            self.i += 1
            return ("op1", tok[1], self.e_unary())
        return self.e_call()

    def e_call(self):
        e = self.primary()
        while True:
            if self.peek() == ("op", ":") and self.peek(1)[0] in ("atom", "var"):
                self.i += 1
                f = self.primary()
                self.expect("op", "(")
                e = ("rcall", e, f, self.exprs_until(")"))
            elif self.peek() == ("op", "("):
                self.i += 1
                e = ("call", e, self.exprs_until(")"))
            else:
                return e

    def primary(self):
        tok = self.next()
        k, v = tok
        if k == "int" or k == "float":
            return ("lit", v)
        if k == "atom":
            return ("lit", atom(v))
        if k == "string":
            return ("lit", from_pylist(v))
        if k == "var":
            return ("var", v)
        if k == "op" and v == "?":                          # macro use: splice its tokens in
            name = self.next()[1]
            if name not in self.macros:
                raise SyntaxError("undefined macro %s" % name)
            self.t[self.i:self.i] = [("op", "(")] + self.macros[name] + [("op", ")")]    # a macro body is one expression
            return self.primary()
        if k == "op" and v == "(":
            e = self.expr()
            self.expect("op", ")")
            return e
        if k == "op" and v == "{":
            return ("tuple", self.exprs_until("}"))
        if k == "op" and v == "[":
            if self.accept("op", "]"):
                return ("lit", NIL)
            first = self.expr()
            if self.accept("op", "||"):                     # list comprehension
                quals = []
                while True:
                    save = self.i
                    pat = self.expr()
                    if self.accept("op", "<-"):
                        quals.append(("gen", pat, self.expr()))
                    else:
                        self.i = save
                        quals.append(("filter", self.expr()))
                    if not self.accept("op", ","):
                        break
                self.expect("op", "]")
                return ("lc", first, quals)
            items = [first]
            while self.accept("op", ","):
                items.append(self.expr())
            tail = ("lit", NIL)
            if self.accept("op", "|"):
                tail = self.expr()
            self.expect("op", "]")
            return ("cons", items, tail)
        if k == "kw" and v == "case":
            subject = self.expr()
            self.expect("kw", "of")
            clauses = []
            while True:
                pat = self.expr()
                guard = self.guard() if self.accept("kw", "when") else None
                self.expect("op", "->")
                clauses.append((pat, guard, self.body()))
                if not self.accept("op", ";"):
                    break
            self.expect("kw", "end")
            return ("case", subject, clauses)
        if k == "kw" and v == "if":
            clauses = []
            while True:
                guard = self.guard()
                self.expect("op", "->")
                clauses.append((guard, self.body()))
                if not self.accept("op", ";"):
                    break
            self.expect("kw", "end")
            return ("if", clauses)
        if k == "kw" and v == "begin":
            b = self.body()
            self.expect("kw", "end")
            return ("block", b)
        if k == "kw" and v == "fun":
            if self.peek() == ("op", "("):
                clauses = []
                while True:
                    clauses.append(self.clause())
                    if not self.accept("op", ";"):
                        break
                self.expect("kw", "end")
                return ("fun", clauses)
            name = self.expect("atom")[1]                   # fun name/arity
            self.expect("op", "/")
            return ("funref", name, self.expect("int")[1])
        raise SyntaxError("unexpected token %r at %d" % (tok, self.i))


# ---------------------------------------------------------------------------------------------------------------------
# evaluator
# ---------------------------------------------------------------------------------------------------------------------
class TailCall:
    __slots__ = ("mod", "name", "args", "fun")

    def __init__(self, mod, name, args, fun=None):
        self.mod, self.name, self.args, self.fun = mod, name, args, fun


def _bool(x) -> bool:
    if x is TRUE:
        return True
    if x is FALSE:
        return False
    raise ErlError("badarg: not a boolean %r" % (x,))


class Machine:
    def __init__(self, src_dir: str, inc_dir: str):
        self.src_dir, self.inc_dir = src_dir, inc_dir
        self.mods: Dict[str, Dict[Tuple[str, int], list]] = {}
        self.devices: List[Device] = []
        self.steps = 0

    def load_text(self, mod: str, text: str):
        """a module from source text (the fs.erl that seeds:generate_fs/3 writes)"""
        self.mods[mod] = Parser(tokenize(text), {"MODULE": [("atom", mod)]}).forms()

    def load(self, mod: str):
        if mod in self.mods:
            return
        path = os.path.join(self.src_dir, mod + ".erl")
        with open(path, encoding="latin-1") as f:
            text = f.read()
        macros: Dict[str, list] = {"MODULE": [("atom", mod)]}
        # -include("x.hrl"): its -define forms come first
        for inc in re.findall(r'^-include\("([^"]+)"\)\.', text, re.M):
            with open(os.path.join(self.inc_dir, os.path.basename(inc)), encoding="latin-1") as f:
                Parser(tokenize(f.read()), macros).forms()
        self.mods[mod] = Parser(tokenize(text), macros).forms()

    # -- calls ----------------------------------------------------------------------------------------------------
    def call(self, mod: str, name: str, args: list):
        """apply with proper tail calls"""
        fun: Optional[Fun] = None
        while True:
            if fun is not None:
                clauses, base_env, cmod = fun.clauses, fun.env, fun.mod
            else:
                if mod not in self.mods:
                    if os.path.exists(os.path.join(self.src_dir, mod + ".erl")):
                        self.load(mod)
                    else:
                        return self.bif(mod, name, args)
                key = (name, len(args))
                if key not in self.mods[mod]:
                    return self.bif(mod, name, args)
                clauses, base_env, cmod = self.mods[mod][key], None, mod
            res = None
            for pats, guard, body in clauses:
                if len(pats) != len(args):
                    continue
                env = dict(base_env) if base_env else {}
                if all(self.match(p, a, env) for p, a in zip(pats, args)) and (guard is None or self.guard(guard, env)):
                    res = self.body(body, env, cmod, True)
                    break
            else:
                raise ErlError("function_clause %s:%s/%d %r" % (mod, name, len(args), args[:2] if fun is None else "fun"))
            if isinstance(res, TailCall):
                mod, name, args, fun = res.mod, res.name, res.args, res.fun
                continue
            return res

    def guard(self, alts, env) -> bool:
        for conj in alts:
            ok = True
            for g in conj:
                try:
                    if self.eval(g, env, None, False) is not TRUE:
                        ok = False
                        break
                except ErlError:
                    ok = False
                    break
            if ok:
                return True
        return False

    def body(self, exprs, env, mod, tail):
        for e in exprs[:-1]:
            self.eval(e, env, mod, False)
        return self.eval(exprs[-1], env, mod, tail)

    # -- patterns -------------------------------------------------------------------------------------------------
    def match(self, pat, val, env) -> bool:
        k = pat[0]
        if k == "var":
            n = pat[1]
            if n == "_":
                return True
            if n in env:
                return exact_eq(env[n], val)
            env[n] = val
            return True
        if k == "lit":
            return exact_eq(pat[1], val) if not isinstance(pat[1], Cons) else (isinstance(val, Cons) and exact_eq(pat[1], val))
        if k == "tuple":
            return isinstance(val, tuple) and len(val) == len(pat[1]) and all(self.match(p, v, env) for p, v in zip(pat[1], val))
        if k == "cons":
            for p in pat[1]:
                if not isinstance(val, Cons) or not self.match(p, val.h, env):
                    return False
                val = val.t
            return self.match(pat[2], val, env)
        if k == "match":                                     # P1 = P2 inside a pattern
            return self.match(pat[1], val, env) and self.match(pat[2], val, env)
        if k == "op2" and pat[1] == "++":                    # "prefix" ++ Rest
            for x in to_pylist(self.eval(pat[2], env, None, False)):
                if not isinstance(val, Cons) or not exact_eq(val.h, x):
                    return False
                val = val.t
            return self.match(pat[3], val, env)
        if k == "op1" and pat[1] == "-" and pat[2][0] == "lit":
            return exact_eq(-pat[2][1], val)
        raise ErlError("unsupported pattern %r" % (pat,))

    # -- expressions ----------------------------------------------------------------------------------------------
    def eval(self, e, env, mod, tail):
        self.steps += 1
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "var":
            try:
                return env[e[1]]
            except KeyError:
                raise ErlError("unbound variable %s" % e[1])
        if k == "tuple":
            return tuple(self.eval(x, env, mod, False) for x in e[1])
        if k == "cons":
            tl = self.eval(e[2], env, mod, False)
            for x in reversed([self.eval(x, env, mod, False) for x in e[1]]):
                tl = Cons(x, tl)
            return tl
        if k == "match":
            v = self.eval(e[2], env, mod, False)
            if not self.match(e[1], v, env):
                raise ErlError("badmatch %r" % (v,))
            return v
        if k == "call":
            f = e[1]
            args = [self.eval(a, env, mod, False) for a in e[2]]
            if f[0] == "lit" and isinstance(f[1], Atom):
                name = str.__str__(f[1])
                if (name, len(args)) in self.mods.get(mod, {}):
                    return TailCall(mod, name, args) if tail else self.call(mod, name, args)
                return self.bif("erlang", name, args)
            fv = self.eval(f, env, mod, False)
            if not isinstance(fv, Fun):
                raise ErlError("badfun %r" % (fv,))
            return TailCall(None, None, args, fv) if tail else self.apply_fun(fv, args)
        if k == "rcall":
            m = self.eval(e[1], env, mod, False)
            f = self.eval(e[2], env, mod, False)
            args = [self.eval(a, env, mod, False) for a in e[3]]
            ms, fs = str.__str__(m), str.__str__(f)
            if ms in self.mods or os.path.exists(os.path.join(self.src_dir, ms + ".erl")):
                self.load(ms)
                if (fs, len(args)) in self.mods[ms]:
                    return TailCall(ms, fs, args) if tail else self.call(ms, fs, args)
            return self.bif(ms, fs, args)
        if k == "case":
            v = self.eval(e[1], env, mod, False)
            for pat, guard, body in e[2]:
                env2 = dict(env)
                if self.match(pat, v, env2) and (guard is None or self.guard(guard, env2)):
                    env.update(env2)
                    return self.body(body, env, mod, tail)
            raise ErlError("case_clause %r" % (v,))
        if k == "if":
            for guard, body in e[1]:
                if self.guard(guard, env):
                    return self.body(body, env, mod, tail)
            raise ErlError("if_clause")
        if k == "block":
            return self.body(e[1], env, mod, tail)
        if k == "fun":
            return Fun(e[1], dict(env), mod)
        if k == "funref":
            name, ar = e[1], e[2]
            params = [("var", "A%d" % i) for i in range(ar)]
            return Fun([(params, None, [("call", ("lit", atom(name)), params)])], {}, mod)
        if k == "lc":
            out: list = []
            self._lc(e[1], e[2], 0, dict(env), mod, out)
            return from_pylist(out)
        if k == "andalso":
            return self.eval(e[2], env, mod, tail) if _bool(self.eval(e[1], env, mod, False)) else FALSE
        if k == "orelse":
            return TRUE if _bool(self.eval(e[1], env, mod, False)) else self.eval(e[2], env, mod, tail)
        if k == "op1":
            v = self.eval(e[2], env, mod, False)
            if e[1] == "-":
                return -v
            if e[1] == "+":
                return v
            if e[1] == "not":
                return FALSE if _bool(v) else TRUE
            if e[1] == "bnot":
                return ~v
        if k == "op2":
            return self.op2(e[1], self.eval(e[2], env, mod, False), self.eval(e[3], env, mod, False))
        raise ErlError("unsupported expression %r" % (e[0],))

    def _lc(self, tmpl, quals, qi, env, mod, out):
        if qi == len(quals):
            out.append(self.eval(tmpl, env, mod, False))
            return
        q = quals[qi]
        if q[0] == "gen":
            for v in to_pylist(self.eval(q[2], env, mod, False)):
                env2 = dict(env)
                if self.match(q[1], v, env2):
                    self._lc(tmpl, quals, qi + 1, env2, mod, out)
        elif _bool(self.eval(q[1], env, mod, False)):
            self._lc(tmpl, quals, qi + 1, env, mod, out)

    def apply_fun(self, f: Fun, args: list):
        return self._run(TailCall(None, None, args, f))

    def _run(self, tc: TailCall):
        # run a fun (or a tail call chain starting with one) to completion
        fun, mod, name, args = tc.fun, tc.mod, tc.name, tc.args
        while True:
            if fun is not None:
                res = None
                for pats, guard, body in fun.clauses:
                    if len(pats) != len(args):
                        continue
                    env = dict(fun.env)
                    if all(self.match(p, a, env) for p, a in zip(pats, args)) and (guard is None or self.guard(guard, env)):
                        res = self.body(body, env, fun.mod, True)
                        break
                else:
                    raise ErlError("function_clause (fun) %r" % (args,))
            else:
                res = self.call(mod, name, args)
            if isinstance(res, TailCall):
                fun, mod, name, args = res.fun, res.mod, res.name, res.args
                continue
            return res

    def op2(self, op, a, b):
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            return a / b
        if op == "div":
            if not isinstance(a, int) or not isinstance(b, int):
                raise ErlError("badarith")
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q          # truncates toward zero
        if op == "rem":
            return a - b * self.op2("div", a, b)
        if op == "bsr":
            return a >> b
        if op == "bsl":
            return a << b
        if op == "band":
            return a & b
        if op == "bor":
            return a | b
        if op == "bxor":
            return a ^ b
        if op == "==":
            return TRUE if compare(a, b) == 0 else FALSE
        if op == "/=":
            return TRUE if compare(a, b) != 0 else FALSE
        if op == "<":
            return TRUE if compare(a, b) < 0 else FALSE
        if op == "=<":
            return TRUE if compare(a, b) <= 0 else FALSE
        if op == ">":
            return TRUE if compare(a, b) > 0 else FALSE
        if op == ">=":
            return TRUE if compare(a, b) >= 0 else FALSE
        if op == "=:=":
            return TRUE if exact_eq(a, b) else FALSE
        if op == "=/=":
            return FALSE if exact_eq(a, b) else TRUE
        if op == "++":
            tl = b
            for x in reversed(to_pylist(a)):
                tl = Cons(x, tl)
            return tl
        if op == "and":
            return TRUE if _bool(a) and _bool(b) else FALSE
        if op == "or":
            return TRUE if _bool(a) or _bool(b) else FALSE
        raise ErlError("unsupported operator %s" % op)

    # -- the OTP functions the four modules use, from their documented semantics -----------------------------------------
    def bif(self, mod, name, args):
        n = len(args)
        if mod in ("erlang", None):
            if name == "length" and n == 1:
                return len(to_pylist(args[0]))
            if name == "element" and n == 2:
                i, t = args
                if not isinstance(i, int) or not isinstance(t, tuple) or i < 1 or i > len(t):
                    raise ErlError("badarg element(%r, tuple of %s)" % (i, len(t) if isinstance(t, tuple) else "?"))
                return t[i - 1]
            if name in ("size", "tuple_size") and n == 1:
                return len(args[0])
            if name == "trunc" and n == 1:
                return int(args[0])
            if name == "integer_to_list" and n == 1:
                return from_pylist(str(args[0]).encode())
            if name == "hd" and n == 1:
                return args[0].h
            if name == "tl" and n == 1:
                return args[0].t
            if name == "max" and n == 2:
                return args[1] if compare(args[0], args[1]) < 0 else args[0]     # erlang:max/2: the first of two equal terms
            if name == "min" and n == 2:
                return args[1] if compare(args[0], args[1]) > 0 else args[0]
            if name == "abs" and n == 1:
                return abs(args[0])
            if name in ("is_list",) and n == 1:
                return TRUE if args[0] is NIL or isinstance(args[0], Cons) else FALSE
            if name == "is_integer" and n == 1:
                return TRUE if isinstance(args[0], int) else FALSE
            if name == "error" and n == 1:
                raise ErlError("error %r" % (args[0],))
        if mod == "lists":
            if name == "reverse" and n == 1:
                out = NIL
                for x in to_pylist(args[0]):
                    out = Cons(x, out)
                return out
            if name == "max" and n == 1:                     # lists:max/1: max([H|T], Max) when H > Max -> H replaces Max
                xs = to_pylist(args[0])
                best = xs[0]
                for x in xs[1:]:
                    if compare(x, best) > 0:
                        best = x
                return best
            if name == "nth" and n == 2:
                return to_pylist(args[1])[args[0] - 1]
            if name == "nthtail" and n == 2:
                l = args[1]
                for _ in range(args[0]):
                    if not isinstance(l, Cons):
                        raise ErlError("function_clause lists:nthtail")
                    l = l.t
                return l
            if name == "seq" and n == 2:
                return from_pylist(range(args[0], args[1] + 1))
            if name == "sort" and n == 1:                    # ascending in term order (stable: equal terms are indistinguishable here)
                import functools
                return from_pylist(sorted(to_pylist(args[0]), key=functools.cmp_to_key(compare)))
            if name == "sublist" and n == 2:
                return from_pylist(to_pylist(args[0])[:args[1]])
            if name == "sublist" and n == 3:
                xs = to_pylist(args[0])
                return from_pylist(xs[args[1] - 1:args[1] - 1 + args[2]])
            if name == "foldl" and n == 3:
                acc = args[1]
                for x in to_pylist(args[2]):
                    acc = self.apply_fun(args[0], [x, acc])
                return acc
            if name == "droplast" and n == 1:
                xs = to_pylist(args[0])
                if not xs:
                    raise ErlError("function_clause lists:droplast([])")
                return from_pylist(xs[:-1])
            if name == "duplicate" and n == 2:
                return from_pylist([args[1]] * args[0])
        if mod == "file" and name == "open" and n == 2:
            dev = Device(bytes(to_pylist(args[0])).decode())
            self.devices.append(dev)
            return (OK, dev)
        if mod == "file" and name == "close" and n == 1:
            return OK
        if mod == "file" and name == "read_line" and n == 1 and isinstance(args[0], ReadDevice):
            ln = args[0].read_line()                                    # list mode: {ok, Line incl. its newline} | eof
            return atom("eof") if ln is None else (OK, from_pylist(ln))
        if mod == "io" and name == "format" and n == 3 and isinstance(args[0], Device):
            fmt, fargs, out, i = to_pylist(args[1]), to_pylist(args[2]), [], 0
            while i < len(fmt):
                c = fmt[i]
                if c == 126 and i + 1 < len(fmt):                       # ~p ~s ~n ~~ ~b ~c: what seeds.erl uses (and a little more)
                    d = chr(fmt[i + 1])
                    i += 2
                    if d == "n":
                        out.append("\n")
                    elif d == "~":
                        out.append("~")
                    elif d in ("p", "w", "b"):
                        v = fargs.pop(0)
                        if not isinstance(v, int):
                            raise ErlError("io:format ~%s of a non-integer is outside the evaluated subset" % d)
                        out.append(str(v))
                    elif d == "s":
                        out.append("".join(map(chr, to_pylist(fargs.pop(0)))))
                    elif d == "c":
                        out.append(chr(fargs.pop(0)))
                    else:
                        raise ErlError("io:format directive ~%s is outside the evaluated subset" % d)
                else:
                    out.append(chr(c))
                    i += 1
            args[0].buf.append("".join(out))
            return OK
        if mod == "io" and name == "format":
            return OK
        raise ErlError("undef %s:%s/%d (not part of the evaluated subset)" % (mod, name, n))


def machine_for_reference(root: str) -> Machine:
    m = Machine(os.path.join(root, "src"), os.path.join(root, "include"))
    for mod in ("fmi", "bwa", "sga", "sw"):
        m.load(mod)
    return m
