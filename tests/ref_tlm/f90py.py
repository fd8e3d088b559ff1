"""TEST INFRASTRUCTURE ONLY (golden-vector generation): a small Fortran-90 -> Python transpiler for the subset the reference's
Tapenade-generated dynamics routines use (model_tlmadm/*.F90: explicit-shape / assumed-shape arrays with arbitrary lower bounds, DO / IF /
SELECT CASE, array sections as actual arguments, OPTIONAL + PRESENT, scalar OUT arguments, pointer aliases of derived-type components,
REAL FUNCTIONs).  tests/golden/make_ref_golden.py uses it IN THIS CONTAINER to execute the reference's own source files from
/root/reference on seeded inputs and to commit the inputs' seeds and the outputs as fixtures; nothing at test time, and nothing under
fv3-jedi-linearmodel_b200/, imports it.  No reference source text is stored in the repository: the transpiled code lives only in memory.

Not supported (asserted): GOTO, WHERE, ALLOCATE, derived-type array components indexed mid-chain, sequence association with a shape change.
"""
import keyword
import re
import numpy as np

# ---------------------------------------------------------------------------------------------------------------- run-time support


class FA:
    """Fortran array: numpy data + lower bounds.  Integer subscripts -> element; slice subscripts use Fortran inclusive bounds."""
    __slots__ = ("a", "lo")

    def __init__(self, a, lo):
        self.a = a
        self.lo = tuple(int(x) for x in lo)
        assert a.ndim == len(self.lo), (a.shape, lo)

    @staticmethod
    def alloc(bounds, dtype=np.float64):
        shape = tuple(max(int(h) - int(l) + 1, 0) for l, h in bounds)
        return FA(np.zeros(shape, dtype=dtype), tuple(l for l, _ in bounds))

    def _ix(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        assert len(idx) == len(self.lo), (idx, self.lo)
        out = []
        for d, (i, l) in enumerate(zip(idx, self.lo)):
            if isinstance(i, slice):
                n = self.a.shape[d]
                a = 0 if i.start is None else i.start - l
                b = n if i.stop is None else i.stop - l + 1
                assert 0 <= a and b <= n and i.step is None, ("section out of bounds", idx, self.lo, self.a.shape)
                out.append(slice(a, max(b, a)))
            else:
                assert isinstance(i, (int, np.integer)), ("non-integer subscript", idx)
                k = i - l
                assert 0 <= k < self.a.shape[d], ("subscript out of bounds", idx, self.lo, self.a.shape)
                out.append(k)
        return tuple(out)

    def __getitem__(self, idx):
        return self.a[self._ix(idx)]

    def __setitem__(self, idx, v):
        self.a[self._ix(idx)] = v

    def sec(self, *idx):
        """array section as an actual argument: a view; sliced dimensions get lower bound 1 (re-based by the callee's declaration)"""
        ix = self._ix(idx)
        v = self.a[ix]
        return FA(v, (1,) * v.ndim)


def bind(actual, bounds, name):
    """dummy array <- actual: explicit-shape bounds (lo, hi) must reproduce the actual's extents; assumed-shape (lo, None) only re-bases"""
    if actual is None:
        return None
    if isinstance(actual, np.ndarray):
        actual = FA(actual, (1,) * actual.ndim)
    assert isinstance(actual, FA), ("array dummy %s got a non-array actual" % name, type(actual))
    assert actual.a.ndim == len(bounds), ("rank mismatch for dummy " + name, actual.a.shape, bounds)
    lo = []
    for d, (l, h) in enumerate(bounds):
        l = 1 if l is None else int(l)
        if h is not None and d < len(bounds) - 1:       # the last extent may differ (sequence association keeps the memory layout):
            #                                                  an access beyond the actual's storage is caught by the subscript check
            assert max(int(h) - l + 1, 0) == actual.a.shape[d], ("extent mismatch for dummy %s dim %d" % (name, d + 1), (l, h), actual.a.shape)
        lo.append(l)
    return FA(actual.a, lo)


import threading                      # noqa: E402
_TLS = threading.local()


def stack():
    """Tapenade's checkpoint stack: one per thread (the six-tile drivers run one thread per tile)"""
    if not hasattr(_TLS, "s"):
        _TLS.s = []
    return _TLS.s


def f_push(kind, v):
    stack().append((kind, v))


def f_pop(kind):
    k, v = stack().pop()
    assert k == kind, ("checkpoint stack out of step: pushed %s, popped as %s" % (k, kind))
    return v


class Goto(Exception):
    """forward GOTO to a labelled statement of an enclosing block: raised at the GOTO, caught in front of the label"""


def f_sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def f_real(x, kind=None):
    return float(x) if isinstance(x, (int, np.integer)) else x


def f_div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


RUNTIME = dict(_any=lambda x: bool(np.any(x.a if isinstance(x, FA) else x)), _stack=stack, _push=f_push, _pop=f_pop, _Goto=Goto, FA=FA, _bind=bind, _alloc=FA.alloc, np=np, _sign=f_sign, _div=f_div, _real=f_real, r_grid=8,
               log=np.log, exp=np.exp, sqrt=np.sqrt, sin=np.sin, cos=np.cos, tan=np.tan, atan=np.arctan, asin=np.arcsin, acos=np.arccos,
               tanh=np.tanh, atan2=np.arctan2)

# ---------------------------------------------------------------------------------------------------------------- source preparation

TOK = re.compile(r"""
    (?P<str>'[^']*'|"[^"]*")
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<num>(?:\d+\.(?!(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge)\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
  | (?P<name>[a-z_]\w*(?:\s*%\s*[a-z_]\w*)*)
  | (?P<op>\*\*|==|/=|<=|>=|=>|//|[-+*/(),:=<>])
  | (?P<ws>\s+)
""", re.X)

DOTOPS = {".and.": " and ", ".or.": " or ", ".not.": " not ", ".eqv.": "==", ".neqv.": "!=", ".eq.": "==", ".ne.": "!=", ".lt.": "<",
          ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".true.": "True", ".false.": "False"}
INTRINSICS = {"abs": "abs", "max": "max", "min": "min", "log": "log", "exp": "exp", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
              "atan": "atan", "asin": "asin", "acos": "acos", "tanh": "tanh", "atan2": "atan2", "sign": "_sign", "real": "_real", "dble": "_real",
              "float": "_real", "any": "_any", "int": "int", "nint": "round", "mod": "_mod"}


def tokenize(s):
    out = []
    pos = 0
    while pos < len(s):
        m = TOK.match(s, pos)
        assert m, ("cannot tokenize", s[pos:pos + 40], s)
        pos = m.end()
        k = m.lastgroup
        if k == "name":
            out.append((k, re.sub(r"\s+", "", m.group(k))))
        elif k != "ws":
            out.append((k, m.group(k)))
    return out


def strip_comment(line):
    q = None
    for i, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:i]
    return line


def logical_lines(text, defines=()):
    """cpp conditionals (only #ifdef / #ifndef / #else / #endif on plain names), comments, continuation lines; lower-cased"""
    lines = []
    stack = []
    for raw in text.split("\n"):
        s = raw.strip()
        if s.startswith("#"):
            w = s[1:].split()
            if w[0] in ("ifdef", "ifndef"):
                stack.append((w[1] in defines) == (w[0] == "ifdef"))
            elif w[0] == "if":
                m = re.match(r"^if\s*(!?)\s*defined\s*\(?\s*(\w+)\s*\)?\s*$", s[1:].strip())
                stack.append(bool(m) and ((m.group(2) in defines) != bool(m.group(1))))
            elif w[0] == "else":
                stack[-1] = not stack[-1]
            elif w[0] == "endif":
                stack.pop()
            continue
        if not all(stack):
            continue
        s = strip_comment(raw).rstrip()
        if not s.strip():
            continue
        lines.append(s)
    out = []
    cur = ""
    for s in lines:
        t = s.strip()
        if t.startswith("&"):
            t = t[1:]
        if cur:
            cur += t
        else:
            cur = t
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        low = "".join(p if (p[:1] in "'\"") else p.lower() for p in re.split(r"""('[^']*'|"[^"]*")""", cur))
        for part in split_top(low, ";"):
            if part.strip():
                out.append(part.strip())
        cur = ""
    return out


def split_top(s, sep=","):
    parts, depth, cur, q = [], 0, "", None
    for c in s:
        if q:
            cur += c
            if c == q:
                q = None
            continue
        if c in "'\"":
            q = c
        elif c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        if c == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += c
    parts.append(cur)
    return parts


def pyname(n):
    parts = n.split("%")
    parts = [p + "_" if (keyword.iskeyword(p) or p in ("is", "in", "lambda")) else p for p in parts]
    return ".".join(parts)

# ---------------------------------------------------------------------------------------------------------------- program units


class Unit:
    def __init__(self, kind, name, args, result=None):
        self.kind, self.name, self.args, self.result = kind, name, args, result
        self.decl = {}            # name -> dict(dims=None | [(lo, hi)], optional, parameter, type, pointer)
        self.body = []
        self.assigned = set()
        self.out_scalars = []     # dummy scalars the unit assigns (returned to the caller, in this order)
        self.do_labels = []


DECL = re.compile(r"^(real|integer|logical|character|type|double\s*precision|complex)\b")
UNIT = re.compile(r"^(?:(?:recursive|pure|elemental)\s+)*(?:(real|integer|logical|double\s*precision)(?:\s*\*\s*\d+|\s*\([^)]*\))?\s+)?"
                  r"(subroutine|function)\s+(\w+)\s*(?:\(([^)]*)\))?(?:\s*result\s*\((\w+)\))?$")


def parse_decl(line, unit_decl):
    if "::" in line:
        head, ents = line.split("::", 1)
    else:
        m = re.match(r"^(real|integer|logical|double\s*precision)(\s*\*\s*\d+|\s*\([^)]*\))?\s+(.*)$", line)
        if not m:
            return
        head, ents = m.group(1), m.group(3)
    attrs = [a.strip() for a in split_top(head)]
    typ = re.match(r"^\w+", attrs[0]).group(0)
    dims = None
    info = dict(optional=False, parameter=False, pointer=False, type=typ)
    for a in attrs[1:]:
        if a.startswith("dimension"):
            dims = a[a.index("(") + 1: a.rindex(")")]
        elif a == "optional":
            info["optional"] = True
        elif a == "parameter":
            info["parameter"] = True
        elif a == "pointer":
            info["pointer"] = True
    for e in split_top(ents):
        e = e.strip()
        init = None
        if "=" in e and "=>" not in e:
            eq = [i for i, c in enumerate(e) if c == "=" and paren_depth(e, i) == 0]
            if eq:
                e, init = e[:eq[0]].strip(), e[eq[0] + 1:].strip()
        m = re.match(r"^(\w+)\s*(?:\((.*)\))?\s*(?:\*\s*\d+)?$", e)
        assert m, ("declaration entity", e, line)
        d = dict(info)
        dd = m.group(2) if m.group(2) is not None else dims
        d["dims"] = None if dd is None else [tuple("" if x.strip() == "*" else x.strip() for x in (split_top(p, ":") if ":" in p else ["1", p]))
                                             for p in split_top(dd)]
        d["init"] = init
        unit_decl[m.group(1)] = d


def paren_depth(s, i):
    return s[:i].count("(") - s[:i].count(")")


def parse_module(text, defines=()):
    """-> (module-level declarations, [Unit])"""
    lines = logical_lines(text, defines)
    mod_decl = {}
    units = []
    cur = None
    for ln in lines:
        m = UNIT.match(ln)
        if m and not ln.startswith("end"):
            args = [a.strip() for a in (m.group(4) or "").split(",") if a.strip()]
            cur = Unit(m.group(2), m.group(3), args, result=(m.group(5) or m.group(3)) if m.group(2) == "function" else None)
            if m.group(2) == "function" and m.group(1):
                cur.decl[cur.result] = dict(optional=False, parameter=False, pointer=False, type=m.group(1), dims=None, init=None)
            units.append(cur)
            continue
        if re.match(r"^end\s*(subroutine|function)\b", ln):
            cur = None
            continue
        if cur is None:
            if DECL.match(ln) and "parameter" in ln:
                parse_decl(ln, mod_decl)
            continue
        if re.match(r"^(use|implicit|intrinsic|external|save|private|public|interface|end\s*interface|module\s+procedure)\b", ln):
            continue
        if DECL.match(ln) and ("::" in ln or re.match(r"^(real|integer|logical|double\s*precision)\s*(\*\s*\d+)?\s+[a-z_]", ln)):
            parse_decl(ln, cur.decl)
            continue
        # labelled DO (`do 555 k = 1, kn` ... `555 continue`): rewritten as a block DO closed right after the labelled statement
        m = re.match(r"^do\s+(\d+)\s*,?\s*(\w+\s*=.*)$", ln)
        if m:
            cur.do_labels.append(m.group(1))
            cur.body.append("do " + m.group(2))
            continue
        cur.body.append(ln)
        m = re.match(r"^(\d+)\s+", ln)
        while m and cur.do_labels and cur.do_labels[-1] == m.group(1):
            cur.do_labels.pop()
            cur.body.append("end do")
    return mod_decl, units

# ---------------------------------------------------------------------------------------------------------------- translation


class Translator:
    def __init__(self, unit, mod_consts, all_units):
        self.cur, self.label_at, self.label_pos, self.opens = 0, {}, {}, {}
        self.blocks = []
        self.u = unit
        self.consts = mod_consts
        self.units = all_units           # name -> Unit (every module)
        self.tmp = 0

    def is_array(self, name):
        if "%" in name:
            return None                   # unknown: decided by the presence of subscripts
        d = self.u.decl.get(name)
        return bool(d and (d["dims"] is not None or d.get("pointer")))

    # ---- expressions
    def expr(self, toks, callarg=False):
        """tokens -> python source.  callarg: a bare array name / section stays an FA object (actual argument of a user procedure)"""
        out = []
        i = 0
        toks = self.int_div(toks)
        n = len(toks)
        while i < n:
            k, v = toks[i]
            if k == "raw":
                out.append(v)
                i += 1
                continue
            if k == "name":
                if i + 1 < n and toks[i + 1] == ("op", "("):
                    j = self.match(toks, i + 1)
                    args = self.split_args(toks[i + 2: j])
                    out.append(self.call_or_index(v, args, callarg and i == 0 and j == n - 1))
                    i = j + 1
                    continue
                if self.is_array(v) and not (callarg and n == 1):
                    out.append(pyname(v) + ".a")
                else:
                    out.append(pyname(v))
            elif k == "num":
                v = re.sub(r"_\w+$", "", v).replace("d", "e")
                if re.match(r"^\d+\.$", v):
                    v += "0"
                elif re.match(r"^\d+\.e", v):
                    v = v.replace(".e", ".0e")
                out.append(v)
            elif k == "dotop":
                out.append(DOTOPS[v])
            elif k == "str":
                out.append(repr(v[1:-1]))
            elif k == "op":
                out.append({"/=": "!=", "//": "+"}.get(v, v))
            i += 1
        return "".join(out)

    def int_div(self, toks):
        """a / b with single-token operands -> _div(a, b): Fortran integer division when both are integers at run time"""
        res = []
        i = 0
        isint = lambda t: t[0] == "num" and re.match(r"^\d+$", t[1])
        simple = lambda t: (t[0] == "name" and not self.is_array(t[1])) or isint(t)
        while i < len(toks):
            if (i + 2 < len(toks) + 0 and toks[i + 1] == ("op", "/") and i + 2 < len(toks) and simple(toks[i]) and simple(toks[i + 2])
                    and (i == 0 or toks[i - 1] not in (("op", "*"), ("op", "/"), ("op", "**"), ("op", ")")))
                    and (i + 3 >= len(toks) or toks[i + 3] not in (("op", "**"), ("op", "(")))):
                f = lambda t: pyname(t[1]) if t[0] == "name" else t[1]
                res.append(("raw", "_div(%s, %s)" % (f(toks[i]), f(toks[i + 2]))))
                i += 3
            else:
                res.append(toks[i])
                i += 1
        return res

    def match(self, toks, i):
        depth = 0
        for j in range(i, len(toks)):
            if toks[j] == ("op", "("):
                depth += 1
            elif toks[j] == ("op", ")"):
                depth -= 1
                if depth == 0:
                    return j
        raise AssertionError("unbalanced parentheses")

    def split_args(self, toks):
        args, depth, cur = [], 0, []
        for t in toks:
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            if t == ("op", ",") and depth == 0:
                args.append(cur)
                cur = []
            else:
                cur.append(t)
        if cur or args:
            args.append(cur)
        return args

    def subscript(self, arg):
        depth = 0
        for p, t in enumerate(arg):
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            elif t == ("op", ":") and depth == 0:
                lo = self.expr(arg[:p]) or "None"
                hi = self.expr(arg[p + 1:]) or "None"
                return "slice(%s, %s)" % (lo, hi), True
        return self.expr(arg), False

    def call_or_index(self, name, args, as_object):
        arr = self.is_array(name)
        if arr is None:
            arr = True
        if arr:
            subs = [self.subscript(a) for a in args]
            sec = any(s for _, s in subs)
            txt = ", ".join(t for t, _ in subs)
            if as_object and sec:
                return "%s.sec(%s)" % (pyname(name), txt)
            return "%s[%s]" % (pyname(name), txt)
        if name == "present":
            return "(%s is not None)" % pyname(args[0][0][1])
        if name in ("size",):
            a = pyname(args[0][0][1])
            return "%s.a.shape[%s - 1]" % (a, self.expr(args[1])) if len(args) > 1 else "%s.a.size" % a
        if name in self.units:
            return self.user_call(name, args)[0] + "[0]"
        if name in INTRINSICS:
            return "%s(%s)" % (INTRINSICS[name], ", ".join(self.expr(a) for a in args))
        if name in self.consts:          # external function supplied by the driver (load(extra=...))
            return "%s(%s)" % (pyname(name), ", ".join(self.expr(a, callarg=True) for a in args))
        raise AssertionError("unknown function or undeclared array: %s in %s" % (name, self.u.name))

    def user_call(self, name, args):
        """-> (python call text, [(position, lvalue text)] of scalar OUT arguments to assign back)"""
        callee = self.units[name]
        pa = []
        back = []
        for p, a in enumerate(args):
            kw = None
            if len(a) > 2 and a[0][0] == "name" and a[1] == ("op", "=") and a[2] != ("op", "="):
                kw, a = a[0][1], a[2:]
            txt = self.expr(a, callarg=True)
            dummy = kw or (callee.args[p] if p < len(callee.args) else None)
            if dummy in callee.out_scalars and a and a[0][0] == "name" and (len(a) == 1 or (a[1] == ("op", "(") and self.match(a, 1) == len(a) - 1)):
                back.append((callee.out_scalars.index(dummy), self.expr(a)))
            pa.append((pyname(kw) + "=" if kw else "") + txt)
        return "%s(%s)" % (pyname(name), ", ".join(pa)), back

    # ---- statements
    def stmt(self, ln, ind, out):
        toks = tokenize(ln)
        k0, v0 = toks[0]
        if k0 == "num" and self.cur in self.label_at:           # labelled statement: close the try block opened for it
            ind -= 1
            out.append("    " * ind + "except _Goto as _g:")
            out.append("    " * ind + "    if _g.args[0] != %s:" % v0)
            out.append("    " * ind + "        raise")
            ln = ln[self.char_pos(ln, toks, 1):]
            toks = toks[1:]
            k0, v0 = toks[0]
            if v0 == "continue":
                return ind
        pad = "    " * ind
        if v0 == "goto" or (v0 == "go" and toks[1][1] == "to"):
            lab = toks[-1][1]
            assert lab in self.label_pos and self.label_pos[lab] > self.cur, ("backward GOTO", ln)
            out.append(pad + "raise _Goto(%s)" % lab)
            return ind
        # one-line IF
        if v0 == "if" and toks[1] == ("op", "("):
            j = self.match(toks, 1)
            cond = self.expr(toks[2:j])
            rest = toks[j + 1:]
            if rest == [("name", "then")]:
                out.append(pad + "if %s:" % cond)
                self.blocks.append(None)
                return ind + 1
            out.append(pad + "if %s:" % cond)
            save = self.cur
            self.cur = -10 - save            # the nested statement is not a block opener / label of its own
            self.cur = save
            self.stmt_nested(ln[self.char_pos(ln, toks, j + 1):], ind + 1, out)
            return ind
        if v0 == "else":
            if len(toks) > 1 and toks[1] == ("name", "if"):
                j = self.match(toks, 2)
                out.append("    " * (ind - 1) + "elif %s:" % self.expr(toks[3:j]))
            else:
                out.append("    " * (ind - 1) + "else:")
            return ind
        if v0 == "elseif":
            j = self.match(toks, 1)
            out.append("    " * (ind - 1) + "elif %s:" % self.expr(toks[2:j]))
            return ind
        if v0 in ("endif", "enddo") or (v0 == "end" and len(toks) > 1 and toks[1][1] in ("if", "do", "select")):
            blk = self.blocks.pop()
            if blk is not None:          # counted DO: the loop variable keeps its final value (first value that fails the test)
                out.append("    " * (ind - 1) + "else:")
                out.append(pad + blk)
            return ind - 1
        if v0 == "do":
            if len(toks) == 1:
                out.append(pad + "while True:")
                self.blocks.append(None)
                return ind + 1
            if toks[1] == ("name", "while"):
                j = self.match(toks, 2)
                out.append(pad + "while %s:" % self.expr(toks[3:j]))
                self.blocks.append(None)
                return ind + 1
            assert toks[1][0] == "name" and toks[2] == ("op", "="), ("DO statement", ln)
            var = pyname(toks[1][1])
            parts = self.split_args(toks[3:])
            a, b = self.expr(parts[0]), self.expr(parts[1])
            c = self.expr(parts[2]) if len(parts) == 3 else "1"
            self.tmp += 1
            lo, hi, st = "_lo%d" % self.tmp, "_hi%d" % self.tmp, "_st%d" % self.tmp
            out.append(pad + "%s, %s, %s = %s, %s, %s" % (lo, hi, st, a, b, c))
            out.append(pad + "for %s in range(%s, %s + (1 if %s > 0 else -1), %s):" % (var, lo, hi, st, st))
            self.blocks.append("%s = %s + max((%s - %s + %s) // %s, 0) * %s" % (var, lo, hi, lo, st, st, st))
            self.u.assigned.add(toks[1][1])
            return ind + 1
        if v0 == "select":
            j = self.match(toks, 2)
            self.tmp += 1
            self.sel = ("_sel%d" % self.tmp, True)
            out.append(pad + "%s = %s" % (self.sel[0], self.expr(toks[3:j])))
            out.append(pad + "if False:")
            out.append(pad + "    pass")
            self.blocks.append(None)
            return ind + 1
        if v0 == "case":
            if toks[1] == ("name", "default"):
                out.append("    " * (ind - 1) + "else:")
            else:
                j = self.match(toks, 1)
                vals = [self.expr(a) for a in self.split_args(toks[2:j])]
                out.append("    " * (ind - 1) + "elif %s in (%s,):" % (self.sel[0], ", ".join(vals)))
            return ind
        if v0 == "call":
            name = toks[1][1]
            args = self.split_args(toks[3:self.match(toks, 2)]) if len(toks) > 2 else []
            m = re.match(r"^(push|pop)(control|integer|realarray|real)(\d+b?|_adm)?$", name)
            if m and name not in self.units:          # Tapenade's checkpoint stack (utils/tapenade): one Python list
                a = args[1] if m.group(2) == "control" and not m.group(3) else args[0]
                bare = len(a) == 1 and a[0][0] == "name" and self.is_array(a[0][1])
                kind = repr(m.group(2)[0] + ("A" if bare else "") + ":" + self.u.name.rsplit("_", 1)[0])
                if m.group(1) == "push":
                    out.append(pad + "_push(%s, %s)" % (kind, pyname(a[0][1]) + ".a.copy()" if bare else self.expr(a)))
                elif bare:
                    out.append(pad + "%s.a[...] = _pop(%s)" % (pyname(a[0][1]), kind))
                else:
                    out.append(pad + "%s = _pop(%s)" % (self.expr(a), kind))
                    self.u.assigned.add(a[0][1])
                return ind
            if name not in self.units:
                out.append(pad + "%s(%s)" % (pyname(name), ", ".join(self.expr(a, callarg=True) for a in args)))     # external stub
                return ind
            txt, back = self.user_call(name, args)
            if back:
                out.append(pad + "_r = " + txt)
                for pos, lv in back:
                    out.append(pad + "%s = _r[%d]" % (lv, pos))
            else:
                out.append(pad + txt)
            return ind
        if v0 == "return":
            out.append(pad + "return " + self.ret())
            return ind
        if v0 in ("exit", "cycle"):
            out.append(pad + ("break" if v0 == "exit" else "continue"))
            return ind
        if v0 in ("print", "write", "stop", "format"):
            out.append(pad + "pass")
            return ind
        assert v0 not in ("where", "allocate", "deallocate") and k0 != "num", ("unsupported statement", ln)
        # assignment / pointer assignment
        depth = 0
        for p, t in enumerate(toks):
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            elif depth == 0 and t in (("op", "="), ("op", "=>")):
                lhs, rhs = toks[:p], toks[p + 1:]
                if t[1] == "=>":
                    out.append(pad + "%s = %s" % (pyname(lhs[0][1]), self.expr(rhs, callarg=True)))
                    self.u.decl.setdefault(lhs[0][1], dict(dims=None, optional=False, parameter=False, type="real", init=None))["pointer"] = True
                    return ind
                name = lhs[0][1]
                self.u.assigned.add(name)
                # function with scalar OUT arguments on the right-hand side (only as the whole right-hand side)
                if rhs and rhs[0][0] == "name" and rhs[0][1] in self.units and len(rhs) > 1 and self.match(rhs, 1) == len(rhs) - 1:
                    txt, back = self.user_call(rhs[0][1], self.split_args(rhs[2:-1]))
                    out.append(pad + "_r = " + txt)
                    out.append(pad + "%s = _r[0]" % self.lvalue(lhs))
                    for pos, lv in back:
                        out.append(pad + "%s = _r[%d]" % (lv, pos + 1))
                    return ind
                out.append(pad + "%s = %s" % (self.lvalue(lhs), self.expr(rhs)))
                return ind
        raise AssertionError("unrecognised statement: " + ln)

    def stmt_nested(self, ln, ind, out):
        la, self.label_at = self.label_at, {}
        try:
            self.stmt(ln, ind, out)
        finally:
            self.label_at = la

    def scan_labels(self):
        """labels and the block each one sits in: label_at[statement index] = label, opens[opener index] = [labels, last first]"""
        self.label_at, self.label_pos, self.opens = {}, {}, {}
        stack = [-1]
        for i, ln in enumerate(self.u.body):
            m = re.match(r"^(\d+)\s+(.*)$", ln)
            t = ln
            if m:
                self.label_at[i] = m.group(1)
                self.label_pos[m.group(1)] = i
                self.opens.setdefault(stack[-1], []).insert(0, m.group(1))
                t = m.group(2)
            if re.match(r"^(end\s*(if|do|select)|endif|enddo)\b", t):
                stack.pop()
            elif re.match(r"^(else|elseif|case)\b", t):
                stack[-1] = i
            elif re.match(r"^(if\s*\(.*\)\s*then|do\b.*|select\s*case.*)$", t) and not re.match(r"^do\s*$", t) or t == "do":
                stack.append(i)
        used = set(re.findall(r"\bgo\s*to\s+(\d+)", "\n".join(self.u.body)))
        for k in list(self.opens):
            self.opens[k] = [l for l in self.opens[k] if l in used]
        self.label_at = {i: l for i, l in self.label_at.items() if l in used}

    def lvalue(self, lhs):
        if len(lhs) == 1:
            n = lhs[0][1]
            return pyname(n) + ".a[...]" if self.is_array(n) else pyname(n)
        return self.expr(lhs)

    def char_pos(self, ln, toks, ntok):
        """character offset of token number ntok in ln"""
        pos = 0
        for _ in range(ntok):
            m = TOK.match(ln, pos)
            while m.lastgroup == "ws":
                pos = m.end()
                m = TOK.match(ln, pos)
            pos = m.end()
        return pos

    def ret(self):
        u = self.u
        items = ([pyname(u.result)] if u.kind == "function" else []) + [pyname(a) for a in u.out_scalars]
        return "(%s)" % "".join(i + ", " for i in items) if items else "None"

    def translate(self):
        u = self.u
        body = []
        ind = 1
        self.scan_labels()
        for _ in self.opens.get(-1, []):
            body.append("    " * ind + "try:")
            ind += 1
        for i, ln in enumerate(u.body):
            self.cur = i
            before = ind
            ind = self.stmt(ln, ind, body)
            if i in self.opens:
                for _ in self.opens[i]:
                    body.append("    " * ind + "try:")
                    ind += 1
        assert ind == 1, ("unbalanced blocks in " + u.name, ind)
        head = []
        sig = []
        seen_opt = False
        for a in u.args:
            d = u.decl.get(a, {})
            seen_opt = seen_opt or bool(d.get("optional"))
            sig.append(pyname(a) + ("=None" if seen_opt else ""))
        head.append("def %s(%s):" % (pyname(u.name), ", ".join(sig)))
        modvars = sorted(n for n in u.assigned if "%" not in n and n not in u.decl and n not in u.args and n != u.result)
        if modvars:                      # module variables the unit sets (visible to the other units of the module)
            head.append("    global " + ", ".join(pyname(n) for n in modvars))
        # dummies first (their bounds may use other dummies), then local arrays; parameters as plain assignments
        for n, d in u.decl.items():
            if d.get("parameter") and d.get("init") is not None:
                head.append("    %s = %s" % (pyname(n), self.expr(tokenize(d["init"]))))
        for n in u.args:
            d = u.decl.get(n)
            if d and d["dims"] is not None:
                b = ", ".join("(%s, %s)" % (self.expr(tokenize(lo)) if lo else "None", self.expr(tokenize(hi)) if hi else "None") for lo, hi in d["dims"])
                head.append("    %s = _bind(%s, (%s,), %r)" % (pyname(n), pyname(n), b, n))
        for n, d in u.decl.items():
            if n in u.args or d["dims"] is None or d.get("pointer") or d.get("parameter"):
                continue
            if any(not hi for _, hi in d["dims"]):
                continue              # deferred shape (pointer declared without the attribute on this line)
            b = ", ".join("(%s, %s)" % (self.expr(tokenize(lo)), self.expr(tokenize(hi))) for lo, hi in d["dims"])
            dt = {"integer": "np.int64", "logical": "np.bool_"}.get(d["type"], "np.float64")
            head.append("    %s = _alloc((%s,), %s)" % (pyname(n), b, dt))
        for n, d in u.decl.items():          # local scalars start defined (Fortran leaves them undefined; they may be passed as OUT actuals)
            if n not in u.args and d["dims"] is None and not d.get("parameter") and not d.get("pointer") and d["type"] in ("real", "integer", "logical", "double"):
                head.append("    %s = %s" % (pyname(n), {"integer": "0", "logical": "False"}.get(d["type"], "0.0")))
        if u.kind == "function":
            head.append("    %s = 0.0" % pyname(u.result))
        return "\n".join(head + body + ["    return " + self.ret()])


def analyse_outs(units):
    """dummy scalars each unit assigns (directly, or by passing them on to another unit's OUT position); fixed point over the call graph"""
    direct = {}
    for u in units.values():
        names = set()
        for ln in u.body:
            toks = tokenize(ln)
            # strip a leading one-line IF
            while toks and toks[0] == ("name", "if") and len(toks) > 1 and toks[1] == ("op", "("):
                depth = 0
                for j in range(1, len(toks)):
                    depth += toks[j] == ("op", "(")
                    depth -= toks[j] == ("op", ")")
                    if depth == 0:
                        break
                toks = toks[j + 1:]
                if toks == [("name", "then")]:
                    toks = []
            if not toks:
                continue
            if toks[0] == ("name", "do") and len(toks) > 2 and toks[2] == ("op", "="):
                names.add(toks[1][1])
                continue
            if toks[0][1] in ("call", "else", "end", "endif", "enddo", "select", "case", "return", "do"):
                continue
            depth = 0
            for p, t in enumerate(toks):
                depth += t == ("op", "(")
                depth -= t == ("op", ")")
                if depth == 0 and t == ("op", "="):
                    names.add(toks[0][1])
                    break
        direct[u.name] = names
    changed = True
    for u in units.values():
        u.out_scalars = [a for a in u.args if a in direct[u.name] and u.decl.get(a, {}).get("dims") is None
                         and u.decl.get(a, {}).get("type") != "type"]
    while changed:
        changed = False
        for u in units.values():
            for ln in u.body:
                m = re.search(r"\bcall\s+(\w+)\s*\((.*)\)\s*$", ln)
                if not m or m.group(1) not in units:
                    continue
                callee = units[m.group(1)]
                for p, a in enumerate(split_top(m.group(2))):
                    a = a.strip()
                    if p < len(callee.args) and callee.args[p] in callee.out_scalars and a in u.args and a not in u.out_scalars \
                            and u.decl.get(a, {}).get("dims") is None:
                        u.out_scalars.append(a)
                        u.out_scalars.sort(key=u.args.index)
                        changed = True


def load(paths, extra=None, defines=(), only=None, strict=True, skip=()):
    """transpile the given Fortran files; -> dict module-path -> namespace (every unit of every file callable from each namespace).
    extra: names injected into every namespace (module variables, constants of modules outside the tree, stubs of external procedures).
    only: optional set of unit names to translate (the others are skipped: they may use unsupported constructs)."""
    parsed = {}
    units = {}
    for p in paths:
        with open(p) as f:
            mod_decl, us = parse_module(f.read(), defines)
        parsed[p] = (mod_decl, us)
        for u in us:
            if (only is None or u.name in only) and u.name not in skip:
                units.setdefault(u.name, u)
    analyse_outs(units)
    spaces = {}
    for p, (mod_decl, us) in parsed.items():
        ns = dict(RUNTIME)
        ns["_mod"] = lambda a, b: a - b * int(a / b)
        ns.update(extra or {})
        dummy = Unit("subroutine", "_module", [])
        dummy.decl = mod_decl
        tr = Translator(dummy, {}, units)
        for n, d in mod_decl.items():
            if d.get("init") is not None and d["dims"] is None:
                try:
                    exec("%s = %s" % (pyname(n), tr.expr(tokenize(d["init"]))), ns)
                except (NameError, AssertionError):      # a constant of a module outside the tree: a unit that reads it fails by name
                    pass
        spaces[p] = ns
    src = {}
    for p, (mod_decl, us) in parsed.items():
        for u in us:
            if units.get(u.name) is not u:
                continue
            try:
                code = Translator(u, {k for k, v in (extra or {}).items() if callable(v)}, units).translate()
            except AssertionError as e:
                if strict:
                    raise
                src[u.name] = "# not translated: %s" % (e,)
                units.pop(u.name)
                continue
            src[u.name] = code
            exec(compile(code, "<f90py:%s>" % u.name, "exec"), spaces[p])
    fns = {}
    for p, (_, us) in parsed.items():
        for u in us:
            if units.get(u.name) is u and pyname(u.name) in spaces[p]:
                fns[u.name] = spaces[p][pyname(u.name)]
    for ns in spaces.values():
        for n, f in fns.items():
            ns.setdefault(pyname(n), f)
    return spaces, fns, src
