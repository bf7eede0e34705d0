"""CEL subset -> (kind, profile, selector bytecode)  (VERDICT r01 #5).

The reference leaves CEL evaluation to kube-scheduler; the expressions its OWN specs use are a small subset:
  * DeviceClass predicates      device.driver == 'gpu.nvidia.com' && device.attributes['gpu.nvidia.com'].type == 'mig'
                                (deployments/helm/k8s-dra-driver/templates/deviceclass-{gpu,mig}.yaml:10)
  * profile choice              device.attributes['gpu.nvidia.com'].profile == '1g.5gb'
                                (demo/specs/quickstart/gpu-test4.yaml:23-25)
  * GPU selection               ...productName.lowerAscii().matches('^.*a100.*$') && (...index == 0 || ...index == 2 ...)
                                (demo/specs/quickstart/gpu-test6.yaml:23-31)
plus comparisons on the other attributes / capacities GpuInfo.GetDevice publishes (cmd/nvidia-dra-plugin/
deviceinfo.go:102-141): index (int), cudaComputeCapability / driverVersion (version), capacity memory (quantity).
This module parses that subset with a recursive-descent parser and lowers it to what the device evaluates
(spec/ALLOCATION.md §10: a postfix program of at most 8 instructions over interned integer attributes) plus the two
facts the host folds into the ClaimRec itself (kind from `type`, profile from `profile ==`).  Anything outside the subset
raises CelError — it is never silently accepted.  Host-side string work only; nothing here touches the GPU.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

from . import records as R
from . import sharing

DOMAIN = "gpu.nvidia.com"


class CelError(ValueError):
    pass


_TOK = re.compile(r"""\s*(?:(?P<str>'(?:[^'\\]|\\.)*'|"(?:[^"\\]|\\.)*")|(?P<num>\d+)|(?P<id>[A-Za-z_][A-Za-z_0-9]*)|(?P<op>&&|\|\||==|!=|<=|>=|[<>!().\[\],]))""")


def _tokens(src: str):
    pos, out = 0, []
    src = src.strip()
    while pos < len(src):
        m = _TOK.match(src, pos)
        if not m or m.end() == pos:
            raise CelError(f"cannot tokenise at {src[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("str") is not None:
            s = m.group("str")[1:-1]
            out.append(("str", re.sub(r"\\(.)", r"\1", s)))
        elif m.group("num") is not None:
            out.append(("num", int(m.group("num"))))
        elif m.group("id") is not None:
            out.append(("id", m.group("id")))
        else:
            out.append(("op", m.group("op")))
    return out


# AST: ("and"|"or", [children]) | ("not", child) | ("cmp", attr_name, space, op, value, methods) | ("true",)
class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if k is None or (kind and k != kind) or (val is not None and v != val):
            raise CelError(f"expected {val or kind}, found {v!r}")
        self.i += 1
        return v

    def expr(self):
        n = self.and_()
        kids = [n]
        while self.peek() == ("op", "||"):
            self.i += 1
            kids.append(self.and_())
        return kids[0] if len(kids) == 1 else ("or", kids)

    def and_(self):
        kids = [self.unary()]
        while self.peek() == ("op", "&&"):
            self.i += 1
            kids.append(self.unary())
        return kids[0] if len(kids) == 1 else ("and", kids)

    def unary(self):
        if self.peek() == ("op", "!"):
            self.i += 1
            return ("not", self.unary())
        if self.peek() == ("op", "("):
            self.i += 1
            n = self.expr()
            self.take("op", ")")
            return n
        return self.comparison()

    def operand(self):
        k, v = self.peek()
        if k == "str" or k == "num":
            self.i += 1
            return ("lit", v)
        if k == "id" and v in ("quantity", "semver"):
            self.i += 1
            self.take("op", "(")
            s = self.take("str")
            self.take("op", ")")
            return ("lit", (v, s))
        if k == "id" and v == "device":
            self.i += 1
            self.take("op", ".")
            what = self.take("id")
            if what == "driver":
                return ("driver",)
            if what not in ("attributes", "capacity"):
                raise CelError(f"device.{what} is outside the supported subset")
            self.take("op", "[")
            dom = self.take("str")
            self.take("op", "]")
            if dom != DOMAIN:
                raise CelError(f"attribute domain {dom!r} is not {DOMAIN!r}")
            self.take("op", ".")
            name = self.take("id")
            methods = []
            while self.peek() == ("op", "."):
                self.i += 1
                mname = self.take("id")
                self.take("op", "(")
                args = []
                while self.peek() != ("op", ")"):
                    args.append(self.operand())
                    if self.peek() == ("op", ","):
                        self.i += 1
                self.take("op", ")")
                methods.append((mname, args))
            return ("attr", what, name, methods)
        raise CelError(f"unexpected token {v!r}")

    def comparison(self):
        lhs = self.operand()
        k, v = self.peek()
        if k == "op" and v in ("==", "!=", "<", "<=", ">", ">="):
            self.i += 1
            rhs = self.operand()
            return ("cmp", lhs, v, rhs)
        return ("cmp", lhs, None, None)               # boolean-valued call such as .matches('...')


def parse(src: str):
    p = _Parser(_tokens(src))
    n = p.expr()
    if p.i != len(p.t):
        raise CelError(f"trailing input at token {p.t[p.i][1]!r}")
    return n


_CMP = {"==": R.CMP_EQ, "!=": R.CMP_NE, "<": R.CMP_LT, "<=": R.CMP_LE, ">": R.CMP_GT, ">=": R.CMP_GE}


@dataclass
class Lowered:
    kind: int | None = None            # R.KIND_GPU / R.KIND_MIG from `type == '...'`
    profile: str | None = None         # from `profile == '...'`
    program: list = field(default_factory=list)     # postfix items for records.selector(); empty = no device-side selector

    def selector(self):
        return R.selector(*self.program) if self.program else None


def _version_major_minor(s: str):
    m = re.match(r"^(\d+)(?:\.(\d+))?(?:\.(\d+))?", s)
    if not m:
        raise CelError(f"bad version {s!r}")
    return int(m.group(1)), int(m.group(2) or 0), int(m.group(3) or 0)


def _leaf(node, products):
    """One comparison -> ('cmp', attr, cmp, value) | ('fact', name, value) | ('const', bool)."""
    _, lhs, op, rhs = node
    if lhs[0] == "driver":
        if op != "==" or rhs[0] != "lit":
            raise CelError("device.driver supports == only")
        return ("const", rhs[1] == DOMAIN)
    if lhs[0] == "lit" and rhs is not None and rhs[0] in ("attr", "driver"):
        flip = {"<": ">", "<=": ">=", ">": "<", ">=": "<=", "==": "==", "!=": "!="}
        return _leaf(("cmp", rhs, flip[op], lhs), products)
    if lhs[0] != "attr":
        raise CelError("comparison needs a device attribute on one side")
    _, space, name, methods = lhs
    mnames = [m[0] for m in methods]
    if space == "capacity":
        if name != "memory" or mnames != ["compareTo"] or rhs is None or rhs[1] != 0:
            raise CelError("capacity subset: memory.compareTo(quantity('..')) <op> 0")
        q = methods[0][1][0][1]
        mib = sharing.quantity_value(q[1] if isinstance(q, tuple) else q) >> 20
        return ("cmp", R.ATTR_MEMORY_MIB, _CMP[op], mib)
    if name == "type":
        if op != "==" or methods:
            raise CelError("type supports == only")
        return ("fact", "type", rhs[1])
    if name == "profile":
        if op != "==" or methods:
            raise CelError("profile supports == only")
        return ("fact", "profile", rhs[1])
    if name == "index":
        if methods or rhs is None or not isinstance(rhs[1], int):
            raise CelError("index compares with an integer")
        return ("cmp", R.ATTR_INDEX, _CMP[op], rhs[1])
    if name in ("cudaComputeCapability", "driverVersion"):
        if mnames != ["compareTo"] or rhs is None or rhs[1] != 0:
            raise CelError(f"{name} subset: .compareTo(semver('..')) <op> 0")
        v = methods[0][1][0][1]
        ma, mi, pa = _version_major_minor(v[1] if isinstance(v, tuple) else v)
        if name == "cudaComputeCapability":
            return ("cmp", R.ATTR_CC, _CMP[op], (ma << 8) | mi)
        if mi or pa:
            raise CelError("driverVersion is compared by its major number only (the device holds the major)")
        return ("cmp", R.ATTR_DRIVER_MAJOR, _CMP[op], ma)
    if name in ("productName", "brand", "architecture"):
        if name != "productName":
            raise CelError(f"{name} is not carried to the device; select by productName")
        lower = "lowerAscii" in mnames
        names = [(p.lower() if lower else p) for p in products]
        if mnames and mnames[-1] == "matches":
            rx = re.compile(methods[-1][1][0][1])
            hit = [i for i, p in enumerate(names) if rx.search(p)]
        elif op in ("==", "!=") and rhs is not None:
            hit = [i for i, p in enumerate(names) if (p == rhs[1]) == (op == "==")]
        else:
            raise CelError("productName subset: == / != / [.lowerAscii()].matches('regex')")
        if any(i >= 32 for i in hit):
            raise CelError("more than 32 interned product names")
        return ("cmp", R.ATTR_PRODUCT, R.CMP_IN_MASK, sum(1 << i for i in hit))
    raise CelError(f"attribute {name!r} is outside the supported subset")


def lower(src: str, products=()) -> Lowered:
    """products: the host's interned productName table (index = the id stored in dra_gpu_attr.product)."""
    out = Lowered()

    def emit(n, top):
        if n[0] == "and":
            first = True
            for k in n[1]:
                before = len(out.program)
                emit(k, top)
                if len(out.program) > before:
                    if not first:
                        out.program.append("and")
                    first = False
            return
        if n[0] == "or":
            # a disjunction of `attr == small int` on ONE attribute folds into a single IN_MASK instruction
            leaves = [_leaf(k, products) if k[0] == "cmp" else None for k in n[1]]
            if all(l and l[0] == "cmp" and l[2] in (R.CMP_EQ, R.CMP_IN_MASK) for l in leaves) and len({l[1] for l in leaves}) == 1 \
                    and all(l[2] == R.CMP_IN_MASK or l[3] < 32 for l in leaves):
                mask = 0
                for l in leaves:
                    mask |= l[3] if l[2] == R.CMP_IN_MASK else (1 << l[3])
                out.program.append(("cmp", leaves[0][1], R.CMP_IN_MASK, mask))
                return
            for i, k in enumerate(n[1]):
                emit(k, False)
                if i:
                    out.program.append("or")
            return
        if n[0] == "not":
            emit(n[1], False)
            out.program.append("not")
            return
        l = _leaf(n, products)
        if l[0] == "fact":
            if not top:
                raise CelError(f"`{l[1]} == ...` must be a top-level conjunct (it is folded into the ClaimRec)")
            if l[1] == "type":
                if l[2] not in ("gpu", "mig"):
                    raise CelError(f"device type {l[2]!r} is not allocated by this path")
                out.kind = R.KIND_GPU if l[2] == "gpu" else R.KIND_MIG
            else:
                out.profile = l[2]
        elif l[0] == "const":
            if not top or not l[1]:
                raise CelError("device.driver must be a top-level conjunct naming this driver")
        else:
            out.program.append(l)

    emit(parse(src), True)
    if len(out.program) > R.SEL_MAX_INS:
        raise CelError(f"selector needs {len(out.program)} instructions, the device evaluates at most {R.SEL_MAX_INS}")
    return out
