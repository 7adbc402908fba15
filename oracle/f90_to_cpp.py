#!/usr/bin/env python
"""f90_to_cpp.py — mechanical Fortran-90-subset -> C++ translator (TEST INFRASTRUCTURE, not product).

Purpose: there is no Fortran compiler in this image, so the reference (`/root/reference/src/greb.f90`,
`src/greb.original.model.f90`) cannot be built with its own Makefile.  This script translates the
reference's modules and subroutines *from the source text where it lies* into C++ that g++ compiles
into `oracle/_ref/lib*.so` (oracle/Makefile target `ref`).  The translated library executes the
reference's own statements — every expression keeps the parse tree the Fortran grammar gives it
(operator precedence, left-to-right association, parentheses), all arithmetic is IEEE binary32 with
no contraction (`-ffp-contract=off`), integer division truncates, `nint` rounds half away from
zero, `x**2` / `x**4` expand to multiplications the way gfortran expands `__builtin_powi`, and the
libm calls (`expf logf cosf sqrtf`) are the same glibc functions gfortran's runtime calls.  It is
what pins the hand-written oracle (oracle/greb_oracle.c): tests compare the two bit for bit.

Nothing of the reference is copied into the repository: the generated C++ and the .so live only in
the git-ignored `oracle/_ref/`.

Supported subset (everything the reference's modules + subroutines use; anything else raises):
modules with scalar/array/parameter declarations and initialisers, `use [, only:]`, implicit
typing, subroutines with implicit interfaces, whole-array and array-section assignments, `where`
(statement and construct with `elsewhere`), `forall`, `do`, `if`/`else if`/`else`, `call`,
direct-access `write(unit,rec=)`, `print *` (numeric items are recorded), intrinsics
exp log sqrt cos abs max min mod nint int float real sum.  The PROGRAM unit (namelist/file I/O) is
skipped: the host side is mirrored by oracle/ref.py which sets the module variables directly.

Every translated module variable is an exported `extern "C"` global named `f_<name>` and every
subroutine an `extern "C" void f_<name>(...)` taking pointers (Fortran passes by reference).
"""
from __future__ import annotations

import re
import sys

INTRINSIC_ELEMENTAL = {"exp", "log", "sqrt", "cos", "sin", "abs", "max", "min", "mod", "nint", "int", "float", "real"}
INTRINSIC_REDUCE = {"sum"}


class F90Error(Exception):
    pass


# ------------------------------------------------------------------------------------------------
# source -> logical statements
# ------------------------------------------------------------------------------------------------
def strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def lower_outside_strings(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def split_semicolons(s: str):
    parts, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def logical_statements(text: str):
    """[(line_no, statement)] with comments removed, continuations joined, ';' split, lower-cased."""
    stmts = []
    pending, pending_line = "", 0
    for no, raw in enumerate(text.splitlines(), 1):
        line = strip_comment(raw)
        if not line.strip():
            continue
        s = line.strip()
        if pending:
            if s.startswith("&"):
                s = s[1:]
            s = pending + " " + s.strip()
            start = pending_line
        else:
            start = no
        if s.endswith("&"):
            pending, pending_line = s[:-1].rstrip(), start
            continue
        pending = ""
        for part in split_semicolons(lower_outside_strings(s)):
            stmts.append((start, part))
    return stmts


# ------------------------------------------------------------------------------------------------
# expression parser (Fortran precedence)
# ------------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<dotop>\.(?:and|or|not|eq|ne|gt|ge|lt|le|true|false)\.)
  | (?P<real>(?:\d+\.(?!(?:and|or|not|eq|ne|gt|ge|lt|le)\.)\d*|\.\d+)(?:[ed][+-]?\d+)?|\d+[ed][+-]?\d+)
  | (?P<int>\d+)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<op>\*\*|==|/=|>=|<=|//|\(/|/\)|[-+*/()=,:<>%])
""", re.X)


def tokenize(s: str):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise F90Error(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        toks.append((k, v))
    return toks


REL = {".eq.": "==", ".ne.": "/=", ".gt.": ">", ".ge.": ">=", ".lt.": "<", ".le.": "<="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "dotop"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise F90Error(f"expected {val!r}, got {self.peek()} in {self.t}")

    def parse_expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.accept(".or."):
            a = ("bin", "||", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.accept(".and."):
            a = ("bin", "&&", a, self.p_not())
        return a

    def p_not(self):
        if self.accept(".not."):
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        k, v = self.peek()
        v = REL.get(v, v)
        if (k in ("op", "dotop")) and v in ("==", "/=", ">", ">=", "<", "<="):
            self.i += 1
            b = self.p_add()
            return ("bin", "!=" if v == "/=" else v, a, b)
        return a

    def p_add(self):
        # unary +/- have the precedence of the additive operators: -a*b == -(a*b)
        if self.peek() == ("op", "-"):
            self.i += 1
            a = ("un", "-", self.p_mul())
        elif self.peek() == ("op", "+"):
            self.i += 1
            a = self.p_mul()
        else:
            a = self.p_mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.accept("**"):
            b = self.p_pow_rhs()
            return ("pow", a, b)
        return a

    def p_pow_rhs(self):
        if self.peek() == ("op", "-"):
            self.i += 1
            return ("un", "-", self.p_pow_rhs())
        a = self.p_primary()
        if self.accept("**"):
            return ("pow", a, self.p_pow_rhs())
        return a

    def p_primary(self):
        k, v = self.next()
        if k == "int":
            return ("int", v)
        if k == "real":
            return ("real", v)
        if k == "str":
            return ("str", v)
        if k == "dotop" and v in (".true.", ".false."):
            return ("bool", v == ".true.")
        if k == "op" and v == "(":
            e = self.parse_expr()
            self.expect(")")
            return ("paren", e)
        if k == "op" and v == "(/":
            items = []
            while True:
                if self.peek() == ("op", "("):  # implied do: (expr, i=a,b)
                    save = self.i
                    self.i += 1
                    e = self.parse_expr()
                    if self.accept(","):
                        kk, var = self.next()
                        if kk == "name" and self.accept("="):
                            lo = self.parse_expr()
                            self.expect(",")
                            hi = self.parse_expr()
                            self.expect(")")
                            items.append(("implied_do", e, var, lo, hi))
                        else:
                            raise F90Error("unsupported array constructor")
                    else:
                        self.i = save
                        items.append(self.parse_expr())
                else:
                    items.append(self.parse_expr())
                if self.accept("/)"):
                    break
                self.expect(",")
            return ("array", items)
        if k == "name":
            if self.peek() == ("op", "("):
                self.i += 1
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.p_subscript())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return ("call", v, args)
            return ("name", v)
        raise F90Error(f"unexpected token {(k, v)} in {self.t}")

    def p_subscript(self):
        # expr | [expr] : [expr] | keyword=expr
        if self.peek() == ("op", ":"):
            self.i += 1
            hi = None
            if self.peek()[1] not in (",", ")"):
                hi = self.parse_expr()
            return ("range", None, hi)
        if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1] == ("op", "=") :
            kw = self.next()[1]
            self.i += 1
            return ("kw", kw, self.parse_expr())
        e = self.parse_expr()
        if self.accept(":"):
            hi = None
            if self.peek()[1] not in (",", ")"):
                hi = self.parse_expr()
            return ("range", e, hi)
        return e


def parse_expression(s: str):
    p = Parser(tokenize(s))
    e = p.parse_expr()
    if p.peek()[0] != "eof":
        raise F90Error(f"trailing tokens in expression {s!r}: {p.t[p.i:]}")
    return e


# ------------------------------------------------------------------------------------------------
# program structure
# ------------------------------------------------------------------------------------------------
class Var:
    def __init__(self, name, typ, dims=None, param=False, init=None, dummy=False, allocatable=False, save=False):
        self.name, self.typ, self.dims = name, typ, dims or []
        self.param, self.init, self.dummy, self.allocatable, self.save = param, init, dummy, allocatable, save

    @property
    def rank(self):
        return len(self.dims)


class Module:
    def __init__(self, name):
        self.name = name
        self.vars = {}      # ordered
        self.uses = []      # [(module, only or None)]


class Sub:
    def __init__(self, name, args, line):
        self.name, self.args, self.line = name, args, line
        self.uses = []
        self.implicit_none = False
        self.locals = {}
        self.body = []      # [(line, text)]


def split_top(s: str, sep=","):
    parts, cur, depth, q = [], [], 0, None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur).strip())
    return parts


DECL_RE = re.compile(r"^(real|integer|logical|character)\b")


def parse_decl(stmt: str):
    """-> list of Var (without dummy flag), or None for character declarations."""
    m = DECL_RE.match(stmt)
    typ = m.group(1)
    rest = stmt[m.end():].strip()
    if typ == "character":
        names = rest.split("::", 1)[1] if "::" in rest else rest
        out = []
        for item in split_top(names.strip()):
            mm = re.match(r"^([a-z_][a-z0-9_]*)", item.strip())
            out.append(Var(mm.group(1), "char"))
        return out
    if rest.startswith("(") :  # real(kind) — not used by the reference
        raise F90Error(f"kind selectors unsupported: {stmt}")
    attrs, names = "", rest
    if "::" in rest:
        attrs, names = rest.split("::", 1)
    attrs_l = [a.strip() for a in split_top(attrs.lstrip(","))] if attrs.strip() else []
    dims, param, alloc = [], False, False
    for a in attrs_l:
        if not a:
            continue
        if a.startswith("dimension"):
            inner = a[a.index("(") + 1:a.rindex(")")]
            dims = [d.strip() for d in split_top(inner)]
        elif a == "parameter":
            param = True
        elif a == "allocatable":
            alloc = True
        elif a.startswith("intent") or a == "save":
            pass
        else:
            raise F90Error(f"unsupported attribute {a!r} in {stmt!r}")
    out = []
    for item in split_top(names.strip()):
        init = None
        if "=" in item:
            # first '=' at depth 0
            depth = 0
            for idx, ch in enumerate(item):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "=" and depth == 0:
                    init = item[idx + 1:].strip()
                    item = item[:idx].strip()
                    break
        vdims = list(dims)
        mm = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", item)
        if not mm:
            raise F90Error(f"cannot parse declarator {item!r} in {stmt!r}")
        if mm.group(3) is not None:
            vdims = [d.strip() for d in split_top(mm.group(3))]
        out.append(Var(mm.group(1), "int" if typ == "integer" else ("bool" if typ == "logical" else "real"),
                       vdims, param, init, allocatable=alloc, save=init is not None))
    return out


def parse_parameter_stmt(stmt: str, known: dict):
    """PARAMETER (name = expr, ...): named constants, implicitly typed unless declared before."""
    inner = stmt[stmt.index("(") + 1:stmt.rindex(")")]
    out = []
    for item in split_top(inner):
        name, init = (x.strip() for x in item.split("=", 1))
        if name in known:
            v = known[name]
            v.param, v.init = True, init
        else:
            v = Var(name, "int" if name[0] in "ijklmn" else "real", [], True, init)
        out.append(v)
    return out


def parse_use(stmt: str):
    m = re.match(r"^use\s+([a-z_][a-z0-9_]*)\s*(,\s*only\s*:\s*(.*))?$", stmt)
    if not m:
        raise F90Error(f"cannot parse {stmt!r}")
    only = None
    if m.group(3) is not None:
        only = [x.strip() for x in m.group(3).split(",") if x.strip()]
    return m.group(1), only


def parse_units(text: str):
    stmts = logical_statements(text)
    modules, subs = {}, {}
    cur, kind = None, None
    for no, s in stmts:
        if kind is None:
            m = re.match(r"^module\s+([a-z_][a-z0-9_]*)$", s)
            if m:
                cur, kind = Module(m.group(1)), "module"
                continue
            m = re.match(r"^subroutine\s+([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", s)
            if m:
                args = [a.strip() for a in (m.group(3) or "").split(",") if a.strip()]
                cur, kind = Sub(m.group(1), args, no), "sub"
                continue
            if re.match(r"^program\b", s):
                cur, kind = None, "program"
                continue
            raise F90Error(f"line {no}: statement outside a program unit: {s!r}")
        if kind == "program":
            if re.match(r"^end(\s+program.*)?$", s):
                kind = None
            continue
        if kind == "module":
            if re.match(r"^end\s*module", s):
                modules[cur.name] = cur
                cur, kind = None, None
            elif s.startswith("use "):
                cur.uses.append(parse_use(s))
            elif s.startswith("namelist") or s == "implicit none":
                pass
            elif DECL_RE.match(s):
                vs = parse_decl(s)
                for v in vs or []:
                    cur.vars[v.name] = v
            elif re.match(r"^parameter\s*\(", s):
                for v in parse_parameter_stmt(s, cur.vars):
                    cur.vars[v.name] = v
            else:
                raise F90Error(f"line {no}: unsupported module statement {s!r}")
            continue
        if kind == "sub":
            if re.match(r"^end(\s*subroutine.*)?$", s):
                subs[cur.name] = cur
                cur, kind = None, None
            elif s.startswith("use ") and not cur.body:
                cur.uses.append(parse_use(s))
            elif s == "implicit none":
                cur.implicit_none = True
            elif DECL_RE.match(s) and "::" in s or (DECL_RE.match(s) and not cur.body and not re.match(r"^(real|integer)\s*=", s)):
                vs = parse_decl(s)
                for v in vs or []:
                    cur.locals[v.name] = v
            elif re.match(r"^parameter\s*\(", s) and not cur.body:
                for v in parse_parameter_stmt(s, cur.locals):
                    cur.locals[v.name] = v
            else:
                cur.body.append((no, s))
            continue
    return modules, subs


# ------------------------------------------------------------------------------------------------
# code generation
# ------------------------------------------------------------------------------------------------
def cname(n: str) -> str:
    return "f_" + n


class Gen:
    def __init__(self, modules, subs):
        self.modules, self.subs = modules, subs
        self.out = []
        self.tmp_id = 0

    # ---- visibility -------------------------------------------------------------------------
    def module_exports(self, mname, seen=None):
        seen = seen or set()
        if mname in seen:
            return {}
        seen.add(mname)
        mod = self.modules[mname]
        names = dict(mod.vars)
        for um, only in mod.uses:
            ex = self.module_exports(um, seen)
            for k, v in ex.items():
                if only is None or k in only:
                    names.setdefault(k, v)
        return names

    def visible_globals(self, uses):
        vis = {}
        for um, only in uses:
            ex = self.module_exports(um)
            if only is None:
                vis.update(ex)
            else:
                for k in only:
                    if k not in ex:
                        raise F90Error(f"use {um}, only: {k} — not exported")
                    vis[k] = ex[k]
        return vis

    # ---- symbol lookup inside a subroutine -----------------------------------------------------
    def setup_sub(self, sub: Sub):
        self.sub = sub
        self.vis = self.visible_globals(sub.uses)
        self.loc = dict(sub.locals)
        for a in sub.args:
            if a in self.loc:
                self.loc[a].dummy = True
            else:
                if sub.implicit_none:
                    raise F90Error(f"{sub.name}: dummy {a} undeclared under implicit none")
                self.loc[a] = Var(a, self.implicit_type(a), dummy=True)
        self.implicit_locals = {}

    @staticmethod
    def implicit_type(name):
        return "int" if name[0] in "ijklmn" else "real"

    def lookup(self, name) -> Var:
        if name in self.loc:
            return self.loc[name]
        if name in self.vis:
            return self.vis[name]
        if name in self.implicit_locals:
            return self.implicit_locals[name]
        if getattr(self, "_dim_ctx", 0):
            for m in self.modules.values():
                if name in m.vars:
                    return m.vars[name]
        if self.sub.implicit_none:
            raise F90Error(f"{self.sub.name}: {name} undeclared under implicit none")
        v = Var(name, self.implicit_type(name))
        self.implicit_locals[name] = v
        return v

    def is_var(self, name):
        return name in self.loc or name in self.vis or name in self.implicit_locals

    # ---- types ---------------------------------------------------------------------------------
    def etype(self, e):
        k = e[0]
        if k == "int":
            return "int"
        if k == "real":
            return "real"
        if k == "bool":
            return "bool"
        if k == "paren":
            return self.etype(e[1])
        if k == "un":
            return "bool" if e[1] == "!" else self.etype(e[2])
        if k == "bin":
            if e[1] in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return "bool"
            a, b = self.etype(e[2]), self.etype(e[3])
            return "real" if "real" in (a, b) else "int"
        if k == "pow":
            return self.etype(e[1])
        if k == "name":
            return self.lookup(e[1]).typ
        if k == "call":
            n = e[1]
            if self.is_array_name(n):
                return self.lookup(n).typ
            if n in ("nint", "int"):
                return "int"
            if n in ("float", "real", "exp", "log", "sqrt", "cos", "sin"):
                return "real"
            if n in ("max", "min", "mod", "abs", "sum"):
                ts = [self.etype(a) for a in e[2]]
                return "real" if "real" in ts else "int"
            raise F90Error(f"{self.sub.name}: unknown function {n}")
        raise F90Error(f"etype: {e}")

    def is_array_name(self, n):
        if n in self.loc or n in self.vis or n in self.implicit_locals:
            return self.lookup(n).rank > 0
        return False

    # ---- rank / shape of an expression ---------------------------------------------------------
    def dims_c(self, v: Var):
        # array extents are evaluated in the scope of the declaration (module parameters), not of the use
        self._dim_ctx = getattr(self, "_dim_ctx", 0) + 1
        try:
            return [self.scalar_expr(parse_expression(d)) if d != ":" else None for d in v.dims]
        finally:
            self._dim_ctx -= 1

    def free_dims(self, e):
        """list of (extent_c, lo_c) of the free dimensions of an array reference, [] if scalar."""
        k = e[0]
        if k == "name":
            v = self.lookup(e[1])
            if v.rank == 0:
                return []
            return [(d, "1") for d in self.dims_c(v)]
        if k == "call" and self.is_array_name(e[1]):
            v = self.lookup(e[1])
            if len(e[2]) != v.rank:
                raise F90Error(f"{self.sub.name}: rank mismatch in {e[1]}")
            dc = self.dims_c(v)
            out = []
            for sub_, d in zip(e[2], dc):
                if sub_[0] == "range":
                    lo = self.scalar_expr(sub_[1]) if sub_[1] is not None else "1"
                    hi = self.scalar_expr(sub_[2]) if sub_[2] is not None else d
                    ext = d if (sub_[1] is None and sub_[2] is None) else f"(({hi})-({lo})+1)"
                    out.append((ext, lo))
            return out
        return None  # not an array reference

    def rank(self, e):
        k = e[0]
        if k in ("int", "real", "bool", "str"):
            return 0
        if k == "paren":
            return self.rank(e[1])
        if k == "un":
            return self.rank(e[2])
        if k == "bin":
            return max(self.rank(e[2]), self.rank(e[3]))
        if k == "pow":
            return self.rank(e[1])
        if k == "name" or (k == "call" and self.is_array_name(e[1])):
            return len(self.free_dims(e))
        if k == "call":
            if e[1] in INTRINSIC_REDUCE:
                return 0
            return max([self.rank(a) for a in e[2]] + [0])
        raise F90Error(f"rank: {e}")

    # ---- expression emission -------------------------------------------------------------------
    def scalar_expr(self, e):
        return self.expr(e, None)

    def ref(self, v: Var) -> str:
        """C lvalue of a scalar variable."""
        if v.dummy and v.rank == 0:
            return f"(*{cname(v.name)})"
        return cname(v.name)

    def flat_index(self, v: Var, idx_c):
        dc = self.dims_c(v)
        s = idx_c[-1]
        for d, ix in zip(reversed(dc[:-1]), reversed(idx_c[:-1])):
            s = f"({ix})+({d})*({s})"
        return s

    def expr(self, e, lv):
        """lv: list of loop-variable names for the free dimensions (elemental context) or None."""
        k = e[0]
        if k == "int":
            return e[1]
        if k == "real":
            t = e[1].replace("d", "e")
            if "." not in t and "e" not in t:
                t += "."
            if "." not in t:
                t = t.replace("e", ".e")
            return t + "f"
        if k == "bool":
            return "true" if e[1] else "false"
        if k == "paren":
            return "(" + self.expr(e[1], lv) + ")"
        if k == "un":
            return f"({e[1]}({self.expr(e[2], lv)}))"
        if k == "bin":
            a, b = self.expr(e[2], lv), self.expr(e[3], lv)
            return f"(({a}){e[1]}({b}))"
        if k == "pow":
            base = self.expr(e[1], lv)
            ex = e[2]
            while ex[0] == "paren":
                ex = ex[1]
            if ex[0] != "int":
                raise F90Error(f"{self.sub.name}: only integer literal exponents supported: {e}")
            n = int(ex[1])
            if self.etype(e[1]) != "real":
                raise F90Error("integer ** unsupported")
            if n not in (2, 3, 4):
                raise F90Error(f"x**{n} unsupported")
            return f"f90_pow{n}({base})"
        if k == "name":
            v = self.lookup(e[1])
            if v.rank == 0:
                return self.ref(v)
            if lv is None:
                raise F90Error(f"{self.sub.name}: array {e[1]} in scalar context")
            if v.rank != len(lv):
                raise F90Error(f"{self.sub.name}: rank of {e[1]} ({v.rank}) != statement rank {len(lv)}")
            return f"{cname(v.name)}[{self.flat_index(v, lv)}]"
        if k == "call":
            n = e[1]
            if self.is_array_name(n):
                v = self.lookup(n)
                idx, used = [], 0
                for sub_ in e[2]:
                    if sub_[0] == "range":
                        if lv is None:
                            raise F90Error(f"{self.sub.name}: section of {n} in scalar context")
                        lo = self.scalar_expr(sub_[1]) if sub_[1] is not None else "1"
                        idx.append(f"({lv[used]}+({lo})-1)")
                        used += 1
                    else:
                        idx.append(f"(({self.scalar_expr(sub_)})-1)")
                if lv is not None and used not in (0, len(lv)):
                    raise F90Error(f"{self.sub.name}: section rank of {n} != statement rank")
                return f"{cname(v.name)}[{self.flat_index(v, idx)}]"
            if n in INTRINSIC_REDUCE:
                arg = e[2][0]
                fd = self.free_dims(arg)
                if fd is None or len(fd) == 0:
                    raise F90Error(f"{self.sub.name}: sum() of a general expression unsupported")
                lvs = [f"_r{i}" for i in range(len(fd))]
                elem = self.expr(arg, lvs)
                typ = "float" if self.etype(arg) == "real" else "int"
                loops = "".join(f"for (int {lvn}=0; {lvn}<({ext}); ++{lvn}) " for lvn, (ext, lo) in reversed(list(zip(lvs, fd))))
                return f"([&]{{ {typ} _s=0; {loops} _s = _s + {elem}; return _s; }}())"
            args = [self.expr(a, lv) for a in e[2]]
            if n == "exp":
                return f"expf({args[0]})"
            if n == "log":
                return f"logf({args[0]})"
            if n == "sqrt":
                return f"sqrtf({args[0]})"
            if n == "cos":
                return f"cosf({args[0]})"
            if n == "sin":
                return f"sinf({args[0]})"
            if n == "abs":
                return f"f90_abs({args[0]})"
            if n in ("max", "min"):
                r = args[0]
                for a in args[1:]:
                    r = f"f90_{n}({r},{a})"
                return r
            if n == "mod":
                if self.etype(e) != "int":
                    raise F90Error("real mod unsupported")
                return f"(({args[0]})%({args[1]}))"
            if n == "nint":
                return f"f90_nint({args[0]})"
            if n == "int":
                return f"((int)({args[0]}))"
            if n in ("float", "real"):
                return f"((float)({args[0]}))"
            raise F90Error(f"{self.sub.name}: unknown function {n}")
        raise F90Error(f"expr: {e}")

    # ---- statements ----------------------------------------------------------------------------
    def emit(self, s):
        self.out.append("  " * self.depth + s)

    def loops_open(self, fd):
        lvs = [f"_i{i}" for i in range(len(fd))]
        for lvn, (ext, lo) in reversed(list(zip(lvs, fd))):  # first dimension innermost
            self.emit(f"for (int {lvn}=0; {lvn}<({ext}); ++{lvn}) {{")
            self.depth += 1
        return lvs

    def loops_close(self, n):
        for _ in range(n):
            self.depth -= 1
            self.emit("}")

    def lhs_parse(self, text):
        e = parse_expression(text)
        if e[0] not in ("name", "call"):
            raise F90Error(f"bad assignment target {text!r}")
        return e

    def assignment(self, lhs_e, rhs_e, mask_e=None, negate=False):
        fd = self.free_dims(lhs_e)
        if fd is None:
            raise F90Error(f"{self.sub.name}: bad lhs {lhs_e}")
        lhs_t = self.etype(lhs_e)
        cast = "(int)" if (lhs_t == "int" and self.etype(rhs_e) == "real") else ""
        if not fd:
            if mask_e is not None:
                raise F90Error("scalar where")
            self.emit(f"{self.expr(lhs_e, None)} = {cast}({self.expr(rhs_e, None)});")
            return
        lvs = self.loops_open(fd)
        st = f"{self.expr(lhs_e, lvs)} = {cast}({self.expr(rhs_e, lvs)});"
        if mask_e is not None:
            m = self.expr(mask_e, lvs)
            st = f"if ({'!' if negate else ''}({m})) {st}"
        self.emit(st)
        self.loops_close(len(fd))

    @staticmethod
    def find_assign(s):
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if s[i + 1:i + 2] == "=" or s[i - 1:i] in ("=", "/", "<", ">"):
                    continue
                return i
        return -1

    @staticmethod
    def matching_paren(s, start):
        depth = 0
        for i in range(start, len(s)):
            if s[i] == "(":
                depth += 1
            elif s[i] == ")":
                depth -= 1
                if depth == 0:
                    return i
        raise F90Error(f"unbalanced parentheses in {s!r}")

    def simple_statement(self, no, s, mask=None, negate=False):
        """assignment / call / print / write / open / one-line where / one-line if."""
        if s == "return":
            self.emit("return;")
            return
        if s.startswith("call "):
            self.call_stmt(s[5:].strip())
            return
        if re.match(r"^print\s*\*", s):
            items = split_top(s[s.index("*") + 1:].lstrip(", "))
            vals = []
            for it in items:
                if not it or it[0] in "'\"":
                    continue
                e = parse_expression(it)
                if e[0] == "name" and self.is_var(e[1]) and self.lookup(e[1]).typ == "char":
                    continue
                if self.rank(e) == 0:
                    vals.append(f"(double)({self.scalar_expr(e)})")
            if vals:
                self.emit(f"f90_print_vals({len(vals)}, {', '.join(vals)});")
            return
        if re.match(r"^open\s*\(", s):
            self.emit(f"/* open: handled by the harness */")
            return
        m = re.match(r"^write\s*\(", s)
        if m:
            close = self.matching_paren(s, m.end() - 1)
            ctrl = split_top(s[m.end():close])
            unit = self.scalar_expr(parse_expression(ctrl[0]))
            rec = None
            for c in ctrl[1:]:
                if c.startswith("rec"):
                    rec = self.scalar_expr(parse_expression(c.split("=", 1)[1]))
            if rec is None:
                raise F90Error("only direct-access writes supported")
            item = parse_expression(s[close + 1:].strip())
            fd = self.lhs_like_dims(item)
            self.emit("{")
            self.depth += 1
            n = "*".join(f"({ext})" for ext, lo in fd)
            self.emit(f"static float _wbuf[{self.const_fold_names(n)}];")
            lvs = self.loops_open(fd)
            flat = lvs[-1]
            for (ext, lo), lvn in zip(reversed(fd[:-1]), reversed(lvs[:-1])):
                flat = f"({lvn})+({ext})*({flat})"
            self.emit(f"_wbuf[{flat}] = {self.expr(item, lvs)};")
            self.loops_close(len(fd))
            self.emit(f"f90_write_rec({unit}, {rec}, _wbuf, {n});")
            self.depth -= 1
            self.emit("}")
            return
        m = re.match(r"^where\s*\(", s)
        if m:
            close = self.matching_paren(s, m.end() - 1)
            mask_e = parse_expression(s[m.end():close])
            rest = s[close + 1:].strip()
            if not rest:
                raise F90Error("block where reached simple_statement")
            i = self.find_assign(rest)
            self.assignment(self.lhs_parse(rest[:i]), parse_expression(rest[i + 1:]), mask_e)
            return
        m = re.match(r"^if\s*\(", s)
        if m:
            close = self.matching_paren(s, m.end() - 1)
            cond = self.scalar_expr(parse_expression(s[m.end():close]))
            rest = s[close + 1:].strip()
            self.emit(f"if ({cond}) {{")
            self.depth += 1
            self.simple_statement(no, rest)
            self.depth -= 1
            self.emit("}")
            return
        i = self.find_assign(s)
        if i < 0:
            raise F90Error(f"line {no}: unsupported statement {s!r}")
        self.assignment(self.lhs_parse(s[:i]), parse_expression(s[i + 1:]), mask, negate)

    def lhs_like_dims(self, e):
        """free dims of a general array expression = those of its first array leaf."""
        fd = self.free_dims(e) if e[0] in ("name", "call") else None
        if fd:
            return fd
        for child in e[1:]:
            if isinstance(child, tuple):
                r = self.lhs_like_dims(child)
                if r:
                    return r
            elif isinstance(child, list):
                for c in child:
                    if isinstance(c, tuple):
                        r = self.lhs_like_dims(c)
                        if r:
                            return r
        return None

    def call_stmt(self, text):
        m = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", text)
        name = m.group(1)
        if name not in self.subs:
            raise F90Error(f"{self.sub.name}: call to unknown subroutine {name}")
        callee = self.subs[name]
        actuals = split_top(m.group(3)) if m.group(3) else []
        if len(actuals) != len(callee.args):
            raise F90Error(f"{self.sub.name}: call {name}: argument count mismatch")
        pre, args = [], []
        for a, formal in zip(actuals, callee.args):
            e = parse_expression(a)
            fv = callee.locals.get(formal) or Var(formal, self.implicit_type(formal))
            ftyp = fv.typ
            if e[0] == "name" and self.is_var(e[1]) or (e[0] == "name" and not self.sub.implicit_none):
                v = self.lookup(e[1])
                if v.typ != ftyp:
                    raise F90Error(f"{self.sub.name}: call {name}: type mismatch for {formal} ({v.typ} vs {ftyp})")
                if (v.rank > 0) != (fv.rank > 0):
                    raise F90Error(f"{self.sub.name}: call {name}: rank mismatch for {formal}")
                if v.rank > 0 or v.dummy:
                    args.append(cname(v.name))
                else:
                    args.append("&" + cname(v.name))
            else:
                if self.rank(e) != 0:
                    raise F90Error(f"{self.sub.name}: array expression as actual argument unsupported")
                t = self.etype(e)
                if t != ftyp:
                    raise F90Error(f"{self.sub.name}: call {name}: type mismatch for {formal} ({t} vs {ftyp})")
                self.tmp_id += 1
                tn = f"_t{self.tmp_id}"
                pre.append(f"{'float' if t == 'real' else 'int'} {tn} = {self.scalar_expr(e)};")
                args.append("&" + tn)
        if pre:
            self.emit("{ " + " ".join(pre))
            self.emit(f"  {cname(name)}({', '.join(args)}); }}")
        else:
            self.emit(f"{cname(name)}({', '.join(args)});")

    def body(self, stmts):
        i = 0
        stack = []  # block kinds
        where_mask = []  # (mask_expr, negate)
        while i < len(stmts):
            no, s = stmts[i]
            i += 1
            try:
                if where_mask and not re.match(r"^(elsewhere|end\s*where)", s):
                    mask_e, neg = where_mask[-1]
                    j = self.find_assign(s)
                    self.assignment(self.lhs_parse(s[:j]), parse_expression(s[j + 1:]), mask_e, neg)
                    continue
                m = re.match(r"^do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", s)
                if m:
                    var = self.lookup(m.group(1))
                    parts = split_top(m.group(2))
                    lo, hi = (self.scalar_expr(parse_expression(p)) for p in parts[:2])
                    if len(parts) > 2:
                        raise F90Error("do with stride unsupported")
                    vr = self.ref(var)
                    self.tmp_id += 1
                    hn = f"_hi{self.tmp_id}"
                    self.emit(f"{{ const int {hn} = {hi}; for ({vr} = {lo}; {vr} <= {hn}; ++{vr}) {{")
                    self.depth += 1
                    stack.append("do")
                    continue
                m = re.match(r"^forall\s*\(\s*([a-z_][a-z0-9_]*)\s*=\s*([^:]+):([^)]+)\)\s*$", s)
                if m:
                    var = self.lookup(m.group(1))
                    lo = self.scalar_expr(parse_expression(m.group(2)))
                    hi = self.scalar_expr(parse_expression(m.group(3)))
                    vr = self.ref(var)
                    self.emit(f"{{ for ({vr} = {lo}; {vr} <= {hi}; ++{vr}) {{")
                    self.depth += 1
                    stack.append("do")
                    continue
                if re.match(r"^end\s*(do|forall)$", s):
                    if stack.pop() != "do":
                        raise F90Error("mismatched end do")
                    self.depth -= 1
                    self.emit("} }")
                    continue
                m = re.match(r"^if\s*\(", s)
                if m:
                    close = self.matching_paren(s, m.end() - 1)
                    if s[close + 1:].strip() == "then":
                        cond = self.scalar_expr(parse_expression(s[m.end():close]))
                        self.emit(f"if ({cond}) {{")
                        self.depth += 1
                        stack.append("if")
                        continue
                m = re.match(r"^else\s*if\s*\(", s)
                if m:
                    close = self.matching_paren(s, m.end() - 1)
                    cond = self.scalar_expr(parse_expression(s[m.end():close]))
                    self.depth -= 1
                    self.emit(f"}} else if ({cond}) {{")
                    self.depth += 1
                    continue
                if s == "else":
                    self.depth -= 1
                    self.emit("} else {")
                    self.depth += 1
                    continue
                if re.match(r"^end\s*if$", s):
                    if stack.pop() != "if":
                        raise F90Error("mismatched end if")
                    self.depth -= 1
                    self.emit("}")
                    continue
                m = re.match(r"^where\s*\(", s)
                if m:
                    close = self.matching_paren(s, m.end() - 1)
                    if not s[close + 1:].strip():
                        where_mask.append((parse_expression(s[m.end():close]), False))
                        continue
                if s == "elsewhere":
                    mask_e, _ = where_mask.pop()
                    where_mask.append((mask_e, True))
                    continue
                if re.match(r"^end\s*where$", s):
                    where_mask.pop()
                    continue
                if s in ("return", "continue"):
                    self.emit("return;" if s == "return" else ";")
                    continue
                self.simple_statement(no, s)
            except F90Error as ex:
                raise F90Error(f"{self.sub.name} line {no}: {ex}\n    statement: {s}") from None
        if stack or where_mask:
            raise F90Error(f"{self.sub.name}: unterminated block")

    # ---- units ---------------------------------------------------------------------------------
    def ctype(self, v):
        return {"real": "float", "int": "int", "bool": "bool"}[v.typ]

    def signature(self, sub: Sub):
        ps = []
        for a in sub.args:
            v = sub.locals.get(a) or Var(a, self.implicit_type(a))
            ps.append(f"{self.ctype(v)}* {cname(a)}")
        return f"void {cname(sub.name)}({', '.join(ps)})"

    def gen_globals(self):
        class _S:  # evaluation context for module-level expressions: every module variable visible
            name = "<module>"
            implicit_none = True
            args = []
        self.sub = _S()
        self.loc, self.implicit_locals = {}, {}
        self.vis = {}
        for m in self.modules.values():
            self.vis.update(m.vars)
        self.depth = 0
        init_code = []
        for m in self.modules.values():
            self.emit(f"// ---- module {m.name}")
            for v in m.vars.values():
                if v.typ == "char":
                    continue
                ct = self.ctype(v)
                if v.allocatable:
                    self.emit(f"{ct}* {cname(v.name)} = nullptr;  // allocatable: set by the harness")
                    continue
                if v.rank == 0:
                    if v.param and v.typ == "int":
                        init = self.scalar_expr(parse_expression(v.init))
                        self.emit(f"constexpr int {cname(v.name)}_c = (int)({self.const_fold_names(init)});")
                        self.emit(f"int {cname(v.name)} = {cname(v.name)}_c;")
                    elif v.init is not None:
                        self.emit(f"{ct} {cname(v.name)} = {self.scalar_expr(parse_expression(v.init))};")
                    else:
                        self.emit(f"{ct} {cname(v.name)} = 0;")
                else:
                    n = "*".join(f"({self.const_fold_names(d)})" for d in self.dims_c(v))
                    if v.init is not None:
                        e = parse_expression(v.init)
                        if e[0] != "array" or any(it[0] == "implied_do" for it in e[1]):
                            raise F90Error(f"module array initialiser of {v.name} unsupported")
                        vals = ", ".join(self.scalar_expr(it) for it in e[1])
                        self.emit(f"{ct} {cname(v.name)}[{n}] = {{{vals}}};")
                    else:
                        self.emit(f"{ct} {cname(v.name)}[{n}];")
        return init_code

    def const_fold_names(self, c_expr: str) -> str:
        """inside constant expressions refer to the constexpr twins of integer parameters."""
        def rep(mo):
            n = mo.group(0)
            base = n[2:]
            for m in self.modules.values():
                v = m.vars.get(base)
                if v is not None and v.param and v.typ == "int" and v.rank == 0:
                    return n + "_c"
            return n
        return re.sub(r"\bf_[a-z_][a-z0-9_]*\b", rep, c_expr)

    def gen_sub(self, sub: Sub):
        self.setup_sub(sub)
        self.depth = 0
        body_out_start = len(self.out)
        self.depth = 1
        self.body(sub.body)
        body_lines = self.out[body_out_start:]
        del self.out[body_out_start:]
        self.depth = 0
        self.emit(f"// ---- subroutine {sub.name} (reference line {sub.line})")
        self.emit(self.signature(sub) + " {")
        self.depth = 1
        for v in list(self.loc.values()) + list(self.implicit_locals.values()):
            if v.dummy:
                continue
            ct = self.ctype(v)
            if v.rank == 0:
                init = f" = {self.scalar_expr(parse_expression(v.init))}" if v.init is not None else " = 0"
                self.emit(f"{'static ' if v.save else ''}{ct} {cname(v.name)}{init};")
            else:
                n = "*".join(f"({self.const_fold_names(d)})" for d in self.dims_c(v))
                self.emit(f"static {ct} {cname(v.name)}[{n}];")
                if v.init is not None:
                    e = parse_expression(v.init)
                    if e[0] != "array":
                        raise F90Error(f"initialiser of {v.name} unsupported")
                    pos = 0
                    for it in e[1]:
                        if it[0] == "implied_do":
                            _, ex, var, lo, hi = it
                            if ex != ("name", var):
                                raise F90Error("general implied do unsupported")
                            lo_c, hi_c = self.scalar_expr(lo), self.scalar_expr(hi)
                            self.emit(f"for (int _k = {lo_c}; _k <= {hi_c}; ++_k) {cname(v.name)}[{pos} + _k - ({lo_c})] = _k;")
                        else:
                            self.emit(f"{cname(v.name)}[{pos}] = {self.scalar_expr(it)};")
                            pos += 1
        self.depth = 0
        self.out.extend(body_lines)
        self.emit("}")
        self.emit("")

    def generate(self, only_subs=None):
        self.depth = 0
        self.out.append(PRELUDE)
        self.out.append('extern "C" {')
        self.gen_globals()
        self.out.append("")
        for s in self.subs.values():
            self.out.append(self.signature(s) + ";")
        self.out.append("")
        for s in self.subs.values():
            self.gen_sub(s)
        self.out.append('}  // extern "C"')
        return "\n".join(self.out) + "\n"


PRELUDE = r"""// GENERATED by oracle/f90_to_cpp.py from the reference Fortran source — do not edit, do not commit.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

static inline float f90_pow2(float x) { return x * x; }
static inline float f90_pow3(float x) { return (x * x) * x; }
static inline float f90_pow4(float x) { const float t = x * x; return t * t; }   // __builtin_powi expansion
static inline int   f90_nint(float x) { return (int)lroundf(x); }                // round half away from zero
static inline float f90_max(float a, float b) { return a > b ? a : b; }
static inline float f90_min(float a, float b) { return a < b ? a : b; }
static inline int   f90_max(int a, int b) { return a > b ? a : b; }
static inline int   f90_min(int a, int b) { return a < b ? a : b; }
static inline float f90_abs(float a) { return fabsf(a); }
static inline int   f90_abs(int a) { return a < 0 ? -a : a; }

// direct-access output and console values are recorded in memory for the Python harness
struct F90Rec { int unit, rec; size_t off, n; };
static std::vector<float>  g_out_data;
static std::vector<F90Rec> g_out_recs;
static std::vector<double> g_print_vals;
static std::vector<int>    g_print_counts;
static int g_record_output = 1;

static void f90_write_rec(int unit, int rec, const float* p, size_t n) {
  if (!g_record_output) return;
  g_out_recs.push_back({unit, rec, g_out_data.size(), n});
  g_out_data.insert(g_out_data.end(), p, p + n);
}
static void f90_print_vals(int n, ...) {
  va_list ap;
  va_start(ap, n);
  for (int i = 0; i < n; ++i) g_print_vals.push_back(va_arg(ap, double));
  va_end(ap);
  g_print_counts.push_back(n);
}
extern "C" {
size_t f90_out_nrecs() { return g_out_recs.size(); }
int f90_out_rec_info(size_t i, int* unit, int* rec, size_t* n) {
  if (i >= g_out_recs.size()) return 1;
  *unit = g_out_recs[i].unit; *rec = g_out_recs[i].rec; *n = g_out_recs[i].n; return 0;
}
const float* f90_out_rec_data(size_t i) { return g_out_data.data() + g_out_recs[i].off; }
void f90_out_reset() { g_out_data.clear(); g_out_recs.clear(); g_print_vals.clear(); g_print_counts.clear(); }
void f90_set_record_output(int on) { g_record_output = on; }
size_t f90_print_nvals() { return g_print_vals.size(); }
const double* f90_print_data() { return g_print_vals.data(); }
size_t f90_print_nlines() { return g_print_counts.size(); }
const int* f90_print_counts() { return g_print_counts.data(); }
}
"""


def translate(path: str) -> str:
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        text = fh.read()
    modules, subs = parse_units(text)
    return Gen(modules, subs).generate()


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit("usage: f90_to_cpp.py <reference.f90> <out.cpp>")
    code = translate(sys.argv[1])
    with open(sys.argv[2], "w") as fh:
        fh.write(code)
