#!/usr/bin/env python
"""Generates ndpp_b200/csrc/legendre_fused.inc: the closed forms of calc_int_pn_tablelin
(src/legendre.F90:22-336) as straight-line code with fewer FP64 instructions and the SAME bits.

Source of truth: the expressions of `int_pn_tablelin_plain` in csrc/legendre.cuh, which restate the Fortran text
operation for operation.  Two rewrites are applied, both exact in IEEE arithmetic as long as no intermediate
result is subnormal or overflows (multiplying by a power of two commutes with rounding):

  1. a power-of-two factor of a constant is pulled out of its product chain:
         ((4*fl)*xh)*xl3  ==  4 * ((fl*xh)*xl3),      ((6*fh)*xh2)  ==  2 * ((3*fh)*xh2)
     which exposes common sub-expressions between the orders (fl*xh, (fh+fl)*xh2, (3*fh)*xh2 ...);
  2. a term carrying such a factor is added with one fused multiply-add, whose single rounding is the rounding
     of the original addition because the scaled product is exact:
         s - 4*t  ==  fma(-4, t, s).

Every other operation keeps the reference's order and association.  The proviso is far from the data: with cosines
drawn from [-1, 1] the host check finds no differing bit for line values anywhere in 1e-290 .. 1e290 (mismatches
appear below 1e-290 and above 1e290, where products underflow or overflow), and on the device a segment whose
numerators leave [2^-969, 2^961) is re-evaluated with the reference text anyway (SharedDivisor::valid).  `python scripts/gen_legendre_fused.py --check`
compiles the generated code for the host (gcc, software-exact fma) and compares it bit for bit with the oracle's
plain restatement (oracle/legendre_ref.c) on random and adversarial inputs; tests/test_host_cpu.py runs that check.
"""
from __future__ import annotations

import argparse
import math
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ndpp_b200", "csrc", "legendre.cuh")
OUT = os.path.join(ROOT, "ndpp_b200", "csrc", "legendre_fused.inc")
MAX_L = 11

_tok = re.compile(r"\s*(?:(\d+\.\d*|\d+)|([A-Za-z_]\w*)|(.))")


def _tokenize(s):
    out = []
    for num, name, op in _tok.findall(s):
        if num:
            out.append(("num", float(num)))
        elif name:
            out.append(("id", name))
        elif op.strip():
            out.append(("op", op))
    return out


class _Parser:
    """C expression -> tree with the left-to-right association of the text."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def next(self):
        x = self.peek()
        self.i += 1
        return x

    def expr(self):
        n = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            n = (op, n, self.term())
        return n

    def term(self):
        n = self.factor()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            n = (op, n, self.factor())
        return n

    def factor(self):
        k, v = self.next()
        if k == "num":
            return ("c", v)
        if k == "id":
            return ("c", {"ONE": 1.0, "TWO": 2.0}[v]) if v in ("ONE", "TWO") else ("v", v)
        if (k, v) == ("op", "("):
            n = self.expr()
            assert self.next() == ("op", ")")
            return n
        raise SyntaxError((k, v))


def reference_trees():
    src = open(SRC).read()
    body = src[src.index("int_pn_tablelin_plain"):src.index("// integrals[l] +=")]
    trees = {}
    for m in re.finditer(r"out\.v\[(\d+)\] = (.*?);\n", body, re.S):
        p = _Parser(_tokenize(m.group(2)))
        trees[int(m.group(1))] = _fold(p.expr())
        assert p.peek()[0] == "end"
    assert sorted(trees) == list(range(MAX_L))
    return trees


def _fold(n):  # constant / constant (ONE / 6.0 is one compile-time constant in the Fortran and in the C)
    if n[0] in "cv":
        return n
    a, b = _fold(n[1]), _fold(n[2])
    if a[0] == "c" and b[0] == "c":
        return ("c", {"+": a[1] + b[1], "-": a[1] - b[1], "*": a[1] * b[1], "/": a[1] / b[1]}[n[0]])
    return (n[0], a, b)


def _pow2split(c):
    """c = 2^k * m with m an odd integer, for integers and dyadic fractions with a small odd part; otherwise (0, c)."""
    m, e = math.frexp(abs(c))
    mant, k = int(m * (1 << 53)), e - 53
    while mant % 2 == 0:
        mant //= 2
        k += 1
    if mant >= (1 << 20):      # 1/6, 1/48, 1/384, 1/3072: not dyadic, stay one constant
        return 0, c
    return k, math.copysign(float(mant), c)


def canon(n):
    """tree -> (k, node): value == 2^k * node exactly.  node kinds: ('c', v), ('v', name), ('*', a, b), ('/', a, b),
    ('sum', ((sign, k, node), ...)) summed left to right."""
    t = n[0]
    if t == "c":
        k, m = _pow2split(n[1])
        return k, ("c", m)
    if t == "v":
        return 0, n
    if t == "*":
        ka, a = canon(n[1])
        kb, b = canon(n[2])
        if a == ("c", 1.0):
            return ka + kb, b
        if b == ("c", 1.0):
            return ka + kb, a
        return ka + kb, ("*", a, b)
    if t == "/":
        ka, a = canon(n[1])
        kb, b = canon(n[2])
        return ka - kb, ("/", a, b)
    items = []

    def flat(m, sign):
        if m[0] in ("+", "-"):
            flat(m[1], sign)
            k, a = canon(m[2])
            items.append((sign if m[0] == "+" else -sign, k, a))
        else:
            k, a = canon(m)
            items.append((sign, k, a))

    flat(n, 1)
    kmin = min(k for _, k, _ in items)   # rounding of a sum scales exactly as well
    return kmin, ("sum", tuple((s, k - kmin, a) for s, k, a in items))


def _lit(v):
    return repr(float(v))


class Emitter:
    def __init__(self):
        self.names, self.lines, self.ops = {}, [], 0

    def ref(self, node):
        if node[0] == "c":
            return _lit(node[1])
        if node[0] == "v":
            return node[1]
        if node in self.names:
            return self.names[node]
        if node[0] == "*":
            rhs = f"{self.ref(node[1])} * {self.ref(node[2])}"
        elif node[0] == "/":
            den = node[2]
            assert den == ("sum", ((1, 0, ("v", "xhigh")), (-1, 0, ("v", "xlow")))), den
            rhs = f"NDPP_DIV({self.ref(node[1])})"
            self.ops += 2   # quotient + two corrections on the device
        else:
            rhs = self.sum(node[1])
        name = f"q{len(self.names)}"
        self.names[node] = name
        self.lines.append(f"{name} = {rhs};")
        self.ops += 1
        return name

    def sum(self, items):
        (s0, k0, a0), rest = items[0], list(items[1:])
        first = self.ref(a0)
        if k0 != 0 and rest and rest[0][1] == 0:
            s1, _, a1 = rest.pop(0)   # 2^k0*a0 +- a1 in one fma
            acc = f"NDPP_FMA({_lit(s0 * 2.0 ** k0)}, {first}, {'-' if s1 < 0 else ''}{self.ref(a1)})"
        elif k0 != 0:
            acc = f"({_lit(s0 * 2.0 ** k0)} * {first})"
            self.ops += 1
        else:
            assert s0 > 0
            acc = first
        for i, (s, k, a) in enumerate(rest):
            r = self.ref(a)
            new = f"NDPP_FMA({_lit(s * 2.0 ** k)}, {r}, {acc})" if k != 0 else f"({acc} {'+' if s > 0 else '-'} {r})"
            if i + 1 < len(rest):
                self.ops += 1
            acc = new
        return acc


def generate():
    trees = reference_trees()
    em = Emitter()
    out = ["// GENERATED by scripts/gen_legendre_fused.py from int_pn_tablelin_plain (csrc/legendre.cuh) -- do not edit.",
           "// Closed forms of calc_int_pn_tablelin (src/legendre.F90:22-336) with exact power-of-two scalings pulled out",
           "// of the product chains and folded into fused multiply-adds; bit-identical to the reference text unless an",
           "// intermediate result is subnormal or overflows.  Needs NDPP_FMA(a, b, c), NDPP_DIV(x) [x / (xhigh - xlow)],",
           "// the scalars xlow, xhigh, flow, fhigh, the powers xl2..xl12 / xh2..xh12, the order count L and t[].", ""]
    marks, blocks = [], []
    for l in range(MAX_L):
        # the statements first needed by order l live in its own block: NDPP_DIV has a side effect (the divisor's range
        # tracking), so an unguarded division of a higher order would stay alive when L is smaller
        k, node = canon(trees[l])
        em.lines = []
        r = em.ref(node)
        val = r if k == 0 else f"{_lit(2.0 ** k)} * {r}"
        if k != 0:
            em.ops += 1
        blocks.append([f"if (L > {l}) {{"] + ["    " + x for x in em.lines] + [f"    t[{l}] = {val};", "}"])
        marks.append(em.ops)
    names = sorted(em.names.values(), key=lambda n: int(n[1:]))
    for i in range(0, len(names), 24):
        out.append("double " + ", ".join(names[i:i + 24]) + ";")
    for b in blocks:
        out += b
    out.append("")
    out.append("// FP64 operations up to and including order l (division = 3): " +
               ", ".join(f"l={l}: {m}" for l, m in enumerate(marks)))
    return "\n".join(out) + "\n", marks


_HARNESS = r"""
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
void ref_calc_int_pn_tablelin(int n, double xlow, double xhigh, double flow, double fhigh, double* integrals);
#define NDPP_FMA(a, b, c) fma(a, b, c)
#define NDPP_DIV(x) ((x) / (xhigh - xlow))
static void fused(int L, double xlow, double xhigh, double flow, double fhigh, double* t)
{
    for (int l = 0; l < 11; ++l) t[l] = 0.0;
    if (xhigh - xlow < 1e-14) return;
    const double xl2 = xlow * xlow, xl3 = xl2 * xlow, xl4 = xl2 * xl2, xl5 = xl2 * xl3, xl6 = xl3 * xl3, xl7 = xl3 * xl4,
                 xl8 = xl4 * xl4, xl9 = xl3 * xl6, xl10 = xl5 * xl5, xl11 = xl5 * xl6, xl12 = xl6 * xl6;
    const double xh2 = xhigh * xhigh, xh3 = xh2 * xhigh, xh4 = xh2 * xh2, xh5 = xh2 * xh3, xh6 = xh3 * xh3, xh7 = xh3 * xh4,
                 xh8 = xh4 * xh4, xh9 = xh3 * xh6, xh10 = xh5 * xh5, xh11 = xh5 * xh6, xh12 = xh6 * xh6;
    (void)xl10; (void)xl11; (void)xl12; (void)xh10; (void)xh11; (void)xh12;
#include "legendre_fused.inc"
}
static uint64_t s = 0x9E3779B97F4A7C15ull;
static double u01(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; }
int main(int argc, char** argv)
{
    long n = argc > 1 ? atol(argv[1]) : 2000000, bad = 0, cmp = 0;
    for (long i = 0; i < n; ++i) {
        double xl, xh, fl, fh;
        const int mode = (int)(i % 8);
        xl = -1.0 + 2.0 * u01();
        if (mode == 0) xh = xl + (1.0 - xl) * u01();                    /* any width */
        else if (mode == 1) xh = xl + 1e-3 * u01();                      /* the default mu spacing */
        else if (mode == 2) { xl = -1.0 + 1e-3 * floor(2000 * u01()); xh = xl + 1e-3; }
        else if (mode == 3) xh = xl + 1e-9 * u01();                      /* nearly empty */
        else if (mode == 4) { xl = -1.0; xh = -1.0 + 2.0 * u01(); }
        else if (mode == 5) { xh = 1.0; }
        else if (mode == 6) { xl = 0.0; xh = u01(); }
        else xh = xl + 1e-13 * u01();                                    /* around FP_PRECISION */
        if (xh > 1.0) xh = 1.0;
        /* physical magnitudes, and every third input anywhere in 1e-250 .. 1e250 */
        const double mag = (i % 3 == 0) ? pow(10.0, -250.0 + 500.0 * u01()) : pow(10.0, -12.0 + 24.0 * u01());
        fl = mag * u01(); fh = mag * u01();
        if (i % 17 == 0) fl = 0.0;
        if (i % 19 == 0) fh = 0.0;
        if (i % 23 == 0) fh = fl;
        double a[11], b[11];
        for (int L = 1; L <= 11; L += (i % 5 == 0 ? 1 : 10)) {
            for (int l = 0; l < 11; ++l) a[l] = 0.0;
            ref_calc_int_pn_tablelin(L, xl, xh, fl, fh, a);
            fused(L, xl, xh, fl, fh, b);
            for (int l = 0; l < L; ++l) {
                ++cmp;
                if (memcmp(&a[l], &b[l], 8) != 0 && !(a[l] != a[l] && b[l] != b[l])) {
                    if (bad < 5) printf("MISMATCH l=%d xl=%.17g xh=%.17g fl=%.17g fh=%.17g ref=%.17g fused=%.17g\n", l, xl, xh, fl, fh, a[l], b[l]);
                    ++bad;
                }
            }
        }
    }
    printf("%ld comparisons, %ld mismatches\n", cmp, bad);
    return bad != 0;
}
"""


def check(n=2000000):
    """Host build of the generated code against oracle/legendre_ref.c, bit for bit."""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "h.c"), "w").write(_HARNESS)
        exe = os.path.join(d, "h")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.dirname(OUT), "-I", os.path.join(ROOT, "oracle"),
                               os.path.join(d, "h.c"), os.path.join(ROOT, "oracle", "legendre_ref.c"), "-o", exe, "-lm"])
        r = subprocess.run([exe, str(n)], capture_output=True, text=True)
        return r.returncode, r.stdout.strip()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--n", type=int, default=2000000)
    a = ap.parse_args()
    text, marks = generate()
    if a.check:
        assert open(OUT).read() == text, "legendre_fused.inc is stale: run scripts/gen_legendre_fused.py"
        rc, msg = check(a.n)
        print(msg)
        sys.exit(rc)
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: FP64 operations up to order l: {marks}")
