#!/usr/bin/env python
"""Generate pyspeedy_b200/csrc/fft96_gen.cuh: straight-line device code for the 96-point real FFT pair.

WHAT: the reference transform (fftpack.f90: rfftb1 = radb2,radb4,radb4,radb3 / rfftf1 = radf3,radf4,radf4,radf2,
with the float-valued tpi/taui/sqrt2/hsqt2 constants) is NOT an exact DFT, so the GPU code must
evaluate the same butterfly DAG.  This script executes those passes symbolically (its own transcription of the
algorithm on expression nodes), then
  * prunes exact identities only (x+0, x-0, 0-x, 0*w): bit-identical to the unpruned evaluation,
    which removes the work on the 35 zero-padded inputs of the inverse (fourier.f90:79-81) and, by dead-code
    elimination, the work for the 35 discarded outputs of the forward transform (fourier.f90:116-121);
  * fuses passes (1,2) and (3,4) into two register-resident stages and splits each stage into its independent
    connected components ("items": 6 x 16-point and 8 x 12-point problems), so that different warps can work on
    different items of the same 32 lines (lane = ensemble member) with ONE shared-memory exchange in between;
  * emits one __device__ function per item.  Addresses are compile-time multiples of the lane stride.

Calling convention of the generated functions (s is the shared-memory exchange buffer, already offset by the
lane; LD / ST are functors supplied by the kernel, which is where prologue/epilogue fusion happens):
  inverse: fftb_A<i>(ld, s)   ld(r)    -> Fourier row r (0..61) of this latitude
           fftb_B<i>(s, st)   st(i, v) <- grid point i (0..95)   (kernel applies 1/cos(lat) scaling, fourier.f90:88-92)
  forward: fftf_A<i>(ld, s)   ld(i)    -> grid point i (kernel applies cosgr scaling / products, spectral.f90:229-242)
           fftf_B<i>(s, st)   st(r, v) <- Fourier row r (kernel multiplies by (double)(1.f/96.f) and zeroes row 1,
                                          fourier.f90:113-121)
Twiddles come from __constant__ c_wa[96] (0-based copy of rffti1's wa) and c_fc[4] = {taui, sqrt2, hsqt2, unused},
filled by the host table generator with glibc cos/sin exactly like the reference.
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 96


# ------------------------------------------------------------------------------------------- expression DAG
class G:
    def __init__(self):
        self.nodes = []  # (op, a, b)
        self.memo = {}

    def mk(self, op, a=None, b=None):
        key = (op, a, b)
        if key in self.memo:
            return self.memo[key]
        self.nodes.append(key)
        self.memo[key] = len(self.nodes) - 1
        return len(self.nodes) - 1


g = G()
ZERO = g.mk("zero")


def inp(name):
    return g.mk("in", name)


def cst(name):
    return g.mk("const", name)


def neg(a):
    if a == ZERO:
        return ZERO
    op, x, _ = g.nodes[a]
    if op == "neg":
        return x
    return g.mk("neg", a)


def add(a, b):
    if a == ZERO:
        return b
    if b == ZERO:
        return a
    return g.mk("add", a, b)


def sub(a, b):
    if b == ZERO:
        return a
    if a == ZERO:
        return neg(b)
    return g.mk("sub", a, b)


def mul(a, b):
    if a == ZERO or b == ZERO:
        return ZERO
    return g.mk("mul", a, b)


# ------------------------------------------------------------------------------------- FFTPACK passes (symbolic)
# cc / ch are python lists of node ids, 0-based storage of the Fortran arrays
def WA(i):  # wa(i), 1-based -> c_wa[i-1]
    return cst("c_wa[%d]" % (i - 1))


TAUI, SQRT2, HSQT2 = cst("c_fc[0]"), cst("c_fc[1]"), cst("c_fc[2]")
HALF = cst("0.5")


def radb2(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, q, k: cc[(i - 1) + ido * ((q - 1) + 2 * (k - 1))]
    def S(i, k, q, v): ch[(i - 1) + ido * ((k - 1) + l1 * (q - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    for k in range(1, l1 + 1):
        S(1, k, 1, add(C(1, 1, k), C(ido, 2, k)))
        S(1, k, 2, sub(C(1, 1, k), C(ido, 2, k)))
    if ido >= 2:
        if ido > 2:
            for k in range(1, l1 + 1):
                for i in range(3, ido + 1, 2):
                    ic = ido + 2 - i
                    S(i - 1, k, 1, add(C(i - 1, 1, k), C(ic - 1, 2, k)))
                    tr2 = sub(C(i - 1, 1, k), C(ic - 1, 2, k))
                    S(i, k, 1, sub(C(i, 1, k), C(ic, 2, k)))
                    ti2 = add(C(i, 1, k), C(ic, 2, k))
                    S(i - 1, k, 2, sub(mul(w1(i - 2), tr2), mul(w1(i - 1), ti2)))
                    S(i, k, 2, add(mul(w1(i - 2), ti2), mul(w1(i - 1), tr2)))
        if ido % 2 == 0:
            for k in range(1, l1 + 1):
                S(ido, k, 1, add(C(ido, 1, k), C(ido, 1, k)))
                S(ido, k, 2, neg(add(C(1, 2, k), C(1, 2, k))))
    return ch


def radb3(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, q, k: cc[(i - 1) + ido * ((q - 1) + 3 * (k - 1))]
    def S(i, k, q, v): ch[(i - 1) + ido * ((k - 1) + l1 * (q - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    w2 = lambda i: WA(iw + ido + i - 1)
    mh = lambda x: neg(mul(HALF, x))  # taur*x with taur = -.5 (exact either way)
    for k in range(1, l1 + 1):
        tr2 = add(C(ido, 2, k), C(ido, 2, k))
        cr2 = add(C(1, 1, k), mh(tr2))
        S(1, k, 1, add(C(1, 1, k), tr2))
        ci3 = mul(TAUI, add(C(1, 3, k), C(1, 3, k)))
        S(1, k, 2, sub(cr2, ci3))
        S(1, k, 3, add(cr2, ci3))
    if ido == 1:
        return ch
    for k in range(1, l1 + 1):
        for i in range(3, ido + 1, 2):
            ic = ido + 2 - i
            tr2 = add(C(i - 1, 3, k), C(ic - 1, 2, k))
            cr2 = add(C(i - 1, 1, k), mh(tr2))
            S(i - 1, k, 1, add(C(i - 1, 1, k), tr2))
            ti2 = sub(C(i, 3, k), C(ic, 2, k))
            ci2 = add(C(i, 1, k), mh(ti2))
            S(i, k, 1, add(C(i, 1, k), ti2))
            cr3 = mul(TAUI, sub(C(i - 1, 3, k), C(ic - 1, 2, k)))
            ci3 = mul(TAUI, add(C(i, 3, k), C(ic, 2, k)))
            dr2, dr3, di2, di3 = sub(cr2, ci3), add(cr2, ci3), add(ci2, cr3), sub(ci2, cr3)
            S(i - 1, k, 2, sub(mul(w1(i - 2), dr2), mul(w1(i - 1), di2)))
            S(i, k, 2, add(mul(w1(i - 2), di2), mul(w1(i - 1), dr2)))
            S(i - 1, k, 3, sub(mul(w2(i - 2), dr3), mul(w2(i - 1), di3)))
            S(i, k, 3, add(mul(w2(i - 2), di3), mul(w2(i - 1), dr3)))
    return ch


def radb4(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, q, k: cc[(i - 1) + ido * ((q - 1) + 4 * (k - 1))]
    def S(i, k, q, v): ch[(i - 1) + ido * ((k - 1) + l1 * (q - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    w2 = lambda i: WA(iw + ido + i - 1)
    w3 = lambda i: WA(iw + 2 * ido + i - 1)
    for k in range(1, l1 + 1):
        tr1 = sub(C(1, 1, k), C(ido, 4, k))
        tr2 = add(C(1, 1, k), C(ido, 4, k))
        tr3 = add(C(ido, 2, k), C(ido, 2, k))
        tr4 = add(C(1, 3, k), C(1, 3, k))
        S(1, k, 1, add(tr2, tr3))
        S(1, k, 2, sub(tr1, tr4))
        S(1, k, 3, sub(tr2, tr3))
        S(1, k, 4, add(tr1, tr4))
    if ido >= 2:
        if ido > 2:
            for k in range(1, l1 + 1):
                for i in range(3, ido + 1, 2):
                    ic = ido + 2 - i
                    ti1 = add(C(i, 1, k), C(ic, 4, k))
                    ti2 = sub(C(i, 1, k), C(ic, 4, k))
                    ti3 = sub(C(i, 3, k), C(ic, 2, k))
                    tr4 = add(C(i, 3, k), C(ic, 2, k))
                    tr1 = sub(C(i - 1, 1, k), C(ic - 1, 4, k))
                    tr2 = add(C(i - 1, 1, k), C(ic - 1, 4, k))
                    ti4 = sub(C(i - 1, 3, k), C(ic - 1, 2, k))
                    tr3 = add(C(i - 1, 3, k), C(ic - 1, 2, k))
                    S(i - 1, k, 1, add(tr2, tr3))
                    cr3 = sub(tr2, tr3)
                    S(i, k, 1, add(ti2, ti3))
                    ci3 = sub(ti2, ti3)
                    cr2, cr4, ci2, ci4 = sub(tr1, tr4), add(tr1, tr4), add(ti1, ti4), sub(ti1, ti4)
                    S(i - 1, k, 2, sub(mul(w1(i - 2), cr2), mul(w1(i - 1), ci2)))
                    S(i, k, 2, add(mul(w1(i - 2), ci2), mul(w1(i - 1), cr2)))
                    S(i - 1, k, 3, sub(mul(w2(i - 2), cr3), mul(w2(i - 1), ci3)))
                    S(i, k, 3, add(mul(w2(i - 2), ci3), mul(w2(i - 1), cr3)))
                    S(i - 1, k, 4, sub(mul(w3(i - 2), cr4), mul(w3(i - 1), ci4)))
                    S(i, k, 4, add(mul(w3(i - 2), ci4), mul(w3(i - 1), cr4)))
        if ido % 2 == 0:
            for k in range(1, l1 + 1):
                ti1 = add(C(1, 2, k), C(1, 4, k))
                ti2 = sub(C(1, 4, k), C(1, 2, k))
                tr1 = sub(C(ido, 1, k), C(ido, 3, k))
                tr2 = add(C(ido, 1, k), C(ido, 3, k))
                S(ido, k, 1, add(tr2, tr2))
                S(ido, k, 2, mul(SQRT2, sub(tr1, ti1)))
                S(ido, k, 3, add(ti2, ti2))
                S(ido, k, 4, neg(mul(SQRT2, add(tr1, ti1))))
    return ch


def radf2(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, k, q: cc[(i - 1) + ido * ((k - 1) + l1 * (q - 1))]
    def S(i, q, k, v): ch[(i - 1) + ido * ((q - 1) + 2 * (k - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    for k in range(1, l1 + 1):
        S(1, 1, k, add(C(1, k, 1), C(1, k, 2)))
        S(ido, 2, k, sub(C(1, k, 1), C(1, k, 2)))
    if ido >= 2:
        if ido > 2:
            for k in range(1, l1 + 1):
                for i in range(3, ido + 1, 2):
                    ic = ido + 2 - i
                    tr2 = add(mul(w1(i - 2), C(i - 1, k, 2)), mul(w1(i - 1), C(i, k, 2)))
                    ti2 = sub(mul(w1(i - 2), C(i, k, 2)), mul(w1(i - 1), C(i - 1, k, 2)))
                    S(i, 1, k, add(C(i, k, 1), ti2))
                    S(ic, 2, k, sub(ti2, C(i, k, 1)))
                    S(i - 1, 1, k, add(C(i - 1, k, 1), tr2))
                    S(ic - 1, 2, k, sub(C(i - 1, k, 1), tr2))
        if ido % 2 == 0:
            for k in range(1, l1 + 1):
                S(1, 2, k, neg(C(ido, k, 2)))
                S(ido, 1, k, C(ido, k, 1))
    return ch


def radf3(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, k, q: cc[(i - 1) + ido * ((k - 1) + l1 * (q - 1))]
    def S(i, q, k, v): ch[(i - 1) + ido * ((q - 1) + 3 * (k - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    w2 = lambda i: WA(iw + ido + i - 1)
    mh = lambda x: neg(mul(HALF, x))
    for k in range(1, l1 + 1):
        cr2 = add(C(1, k, 2), C(1, k, 3))
        S(1, 1, k, add(C(1, k, 1), cr2))
        S(1, 3, k, mul(TAUI, sub(C(1, k, 3), C(1, k, 2))))
        S(ido, 2, k, add(C(1, k, 1), mh(cr2)))
    if ido == 1:
        return ch
    for k in range(1, l1 + 1):
        for i in range(3, ido + 1, 2):
            ic = ido + 2 - i
            dr2 = add(mul(w1(i - 2), C(i - 1, k, 2)), mul(w1(i - 1), C(i, k, 2)))
            di2 = sub(mul(w1(i - 2), C(i, k, 2)), mul(w1(i - 1), C(i - 1, k, 2)))
            dr3 = add(mul(w2(i - 2), C(i - 1, k, 3)), mul(w2(i - 1), C(i, k, 3)))
            di3 = sub(mul(w2(i - 2), C(i, k, 3)), mul(w2(i - 1), C(i - 1, k, 3)))
            cr2, ci2 = add(dr2, dr3), add(di2, di3)
            S(i - 1, 1, k, add(C(i - 1, k, 1), cr2))
            S(i, 1, k, add(C(i, k, 1), ci2))
            tr2 = add(C(i - 1, k, 1), mh(cr2))
            ti2 = add(C(i, k, 1), mh(ci2))
            tr3 = mul(TAUI, sub(di2, di3))
            ti3 = mul(TAUI, sub(dr3, dr2))
            S(i - 1, 3, k, add(tr2, tr3))
            S(ic - 1, 2, k, sub(tr2, tr3))
            S(i, 3, k, add(ti2, ti3))
            S(ic, 2, k, sub(ti3, ti2))
    return ch


def radf4(ido, l1, cc, iw):
    ch = [ZERO] * N
    C = lambda i, k, q: cc[(i - 1) + ido * ((k - 1) + l1 * (q - 1))]
    def S(i, q, k, v): ch[(i - 1) + ido * ((q - 1) + 4 * (k - 1))] = v
    w1 = lambda i: WA(iw + i - 1)
    w2 = lambda i: WA(iw + ido + i - 1)
    w3 = lambda i: WA(iw + 2 * ido + i - 1)
    for k in range(1, l1 + 1):
        tr1 = add(C(1, k, 2), C(1, k, 4))
        tr2 = add(C(1, k, 1), C(1, k, 3))
        S(1, 1, k, add(tr1, tr2))
        S(ido, 4, k, sub(tr2, tr1))
        S(ido, 2, k, sub(C(1, k, 1), C(1, k, 3)))
        S(1, 3, k, sub(C(1, k, 4), C(1, k, 2)))
    if ido >= 2:
        if ido > 2:
            for k in range(1, l1 + 1):
                for i in range(3, ido + 1, 2):
                    ic = ido + 2 - i
                    cr2 = add(mul(w1(i - 2), C(i - 1, k, 2)), mul(w1(i - 1), C(i, k, 2)))
                    ci2 = sub(mul(w1(i - 2), C(i, k, 2)), mul(w1(i - 1), C(i - 1, k, 2)))
                    cr3 = add(mul(w2(i - 2), C(i - 1, k, 3)), mul(w2(i - 1), C(i, k, 3)))
                    ci3 = sub(mul(w2(i - 2), C(i, k, 3)), mul(w2(i - 1), C(i - 1, k, 3)))
                    cr4 = add(mul(w3(i - 2), C(i - 1, k, 4)), mul(w3(i - 1), C(i, k, 4)))
                    ci4 = sub(mul(w3(i - 2), C(i, k, 4)), mul(w3(i - 1), C(i - 1, k, 4)))
                    tr1, tr4, ti1, ti4 = add(cr2, cr4), sub(cr4, cr2), add(ci2, ci4), sub(ci2, ci4)
                    ti2, ti3 = add(C(i, k, 1), ci3), sub(C(i, k, 1), ci3)
                    tr2, tr3 = add(C(i - 1, k, 1), cr3), sub(C(i - 1, k, 1), cr3)
                    S(i - 1, 1, k, add(tr1, tr2))
                    S(ic - 1, 4, k, sub(tr2, tr1))
                    S(i, 1, k, add(ti1, ti2))
                    S(ic, 4, k, sub(ti1, ti2))
                    S(i - 1, 3, k, add(ti4, tr3))
                    S(ic - 1, 2, k, sub(tr3, ti4))
                    S(i, 3, k, add(tr4, ti3))
                    S(ic, 2, k, sub(tr4, ti3))
        if ido % 2 == 0:
            for k in range(1, l1 + 1):
                ti1 = neg(mul(HSQT2, add(C(ido, k, 2), C(ido, k, 4))))
                tr1 = mul(HSQT2, sub(C(ido, k, 2), C(ido, k, 4)))
                S(ido, 1, k, add(tr1, C(ido, k, 1)))
                S(ido, 3, k, sub(C(ido, k, 1), tr1))
                S(1, 2, k, sub(ti1, C(ido, k, 3)))
                S(1, 4, k, add(ti1, C(ido, k, 3)))
    return ch


BACK = [(2, radb2), (4, radb4), (4, radb4), (3, radb3)]  # ifac = [96,4,2,4,4,3], fftpack.f90:69-134
FWD = [(3, radf3), (4, radf4), (4, radf4), (2, radf2)]   # fftpack.f90:136-202 (reverse order)


def run_back(x, passes):
    """passes: indices into BACK (0..3) to apply, starting with l1/iw as rfftb1 would have them."""
    l1, iw = 1, 1
    for p, (ip, fn) in enumerate(BACK):
        ido = N // (l1 * ip)
        if p in passes:
            x = fn(ido, l1, x, iw)
        l1 *= ip
        iw += (ip - 1) * ido
    return x


def run_fwd(x, passes):
    l2, iw = N, N
    for p, (ip, fn) in enumerate(FWD):
        l1 = l2 // ip
        ido = N // l2
        iw -= (ip - 1) * ido
        if p in passes:
            x = fn(ido, l1, x, iw)
        l2 = l1
    return x


# ------------------------------------------------------------------------------------------------- analysis
def leaves(node, memo):
    if node in memo:
        return memo[node]
    op, a, b = g.nodes[node]
    if op == "in":
        r = frozenset([node])
    elif op in ("zero", "const"):
        r = frozenset()
    else:
        r = leaves(a, memo)
        if b is not None and op != "neg":
            r = r | leaves(b, memo)
    memo[node] = r
    return r


def components(outs):
    """Group output slots (index -> node) into connected components via shared input leaves."""
    memo = {}
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for slot, node in outs.items():
        key = ("o", slot)
        find(key)
        for lf in leaves(node, memo):
            parent[find(("l", lf))] = find(key)
    groups = defaultdict(list)
    for slot in outs:
        groups[find(("o", slot))].append(slot)
    return sorted((sorted(v) for v in groups.values()), key=lambda v: v[0])


LEVEL_ORDER = False


def emit(fname, args, outs, load_expr, store_stmt):
    """outs: dict slot -> node.  Returns CUDA source of one item function + flop count."""
    order, seen = [], set()

    def visit(n):
        if n in seen:
            return
        seen.add(n)
        op, a, b = g.nodes[n]
        if op in ("add", "sub", "mul"):
            visit(a), visit(b)
        elif op == "neg":
            visit(a)
        order.append(n)

    for slot in sorted(outs):
        visit(outs[slot])
    if LEVEL_ORDER:
        # breadth-first (level) order: consecutive instructions are independent of each other, which matters when a
        # single warp per scheduler has to cover the FP64 latency on its own (whole-line register variant)
        depth = {}
        for n in order:  # `order` is a valid topological order
            op, a, b = g.nodes[n]
            if op in ("add", "sub", "mul"):
                depth[n] = 1 + max(depth[a], depth[b])
            elif op == "neg":
                depth[n] = depth[a]
            else:
                depth[n] = 0
        order.sort(key=lambda n: depth[n])
    name = {}
    lines = []
    flops = 0
    for n in order:
        op, a, b = g.nodes[n]
        if op == "zero":
            name[n] = "0.0"
        elif op == "const":
            name[n] = a
        elif op == "in":
            name[n] = "x%d" % n
            lines.append("    const T x%d = %s;" % (n, load_expr(a)))
        elif op == "neg":
            name[n] = "(-%s)" % name[a]
        else:
            sym = {"add": "+", "sub": "-", "mul": "*"}[op]
            name[n] = "t%d" % n
            lines.append("    const T t%d = %s %s %s;" % (n, name[a], sym, name[b]))
            flops += 1
    for slot in sorted(outs):
        lines.append("    " + store_stmt(slot, name[outs[slot]]))
    # T = value type: double (one line per thread) or a pair of lines (D2, fused_common.cuh: 128-bit shared-memory accesses)
    tmpl = "template <class LD, class T> " if "LD ld" in args else "template <class T, class ST> "
    src = tmpl + "__device__ __forceinline__ void %s(%s) {\n%s\n}\n" % (fname, args, "\n".join(lines))
    return src, flops


def main():
    global LEVEL_ORDER
    out = []
    out.append("// GENERATED by tools/gen_fft96.py -- do not edit.  See that script for the derivation.\n"
               "// 96-point real FFT pair equivalent to the reference's FFTPACK path (fftpack.f90:69-202 with the\n"
               "// N=96 factorisation 2,4,4,3), as zero-pruned straight-line items.  FFT_LS = lane stride (in values of type T:\n"
               "// double, or a pair of lines).\n"
               "#pragma once\n")
    report = []

    # ---------------- inverse: fvar(1)=row0, fvar(m-1)=row m (m=3..62) -> fvar index e>=1 is row e+1; e>=61 zero
    fvar = [ZERO] * N
    fvar[0] = inp(("four", 0))
    for e in range(1, 61):
        fvar[e] = inp(("four", e + 1))
    mid = run_back(fvar, {0, 1})
    comps = components({s: n for s, n in enumerate(mid)})
    report.append("inverse stage A items: " + str([len(c) for c in comps]))
    nA = len(comps)
    for q, slots in enumerate(comps):
        src, fl = emit("fftb_A%d" % q, "const LD ld, T* __restrict__ s",
                       {s: mid[s] for s in slots},
                       lambda a: "ld(%d)" % a[1],
                       lambda slot, v: "s[%d * FFT_LS] = %s;" % (slot, v))
        out.append(src)
        report.append("  A%d: %d outputs, %d flops" % (q, len(slots), fl))
    sin = [inp(("s", e)) for e in range(N)]
    fin = run_back(sin, {2, 3})
    comps = components({s: n for s, n in enumerate(fin)})
    report.append("inverse stage B items: " + str([len(c) for c in comps]))
    nB = len(comps)
    for q, slots in enumerate(comps):
        src, fl = emit("fftb_B%d" % q, "const T* __restrict__ s, const ST st",
                       {s: fin[s] for s in slots},
                       lambda a: "s[%d * FFT_LS]" % a[1],
                       lambda slot, v: "st(%d, %s);" % (slot, v))
        out.append(src)
        report.append("  B%d: %d outputs, %d flops" % (q, len(slots), fl))
    out.append("#define FFTB_NA %d\n#define FFTB_NB %d\n" % (nA, nB))

    # ---------------- forward: outputs kept: fvar(1) -> row 0 ; fvar(m-1) -> row m, m=3..62 (i.e. e=1..60 -> row e+1)
    gin = [inp(("grid", e)) for e in range(N)]
    mid = run_fwd(gin, {0, 1})
    comps = components({s: n for s, n in enumerate(mid)})
    report.append("forward stage A items: " + str([len(c) for c in comps]))
    nA = len(comps)
    for q, slots in enumerate(comps):
        src, fl = emit("fftf_A%d" % q, "const LD ld, T* __restrict__ s",
                       {s: mid[s] for s in slots},
                       lambda a: "ld(%d)" % a[1],
                       lambda slot, v: "s[%d * FFT_LS] = %s;" % (slot, v))
        out.append(src)
        report.append("  A%d: %d outputs, %d flops" % (q, len(slots), fl))
    sin = [inp(("s", e)) for e in range(N)]
    fin = run_fwd(sin, {2, 3})
    keep = {0: fin[0]}
    for e in range(1, 61):
        keep[e + 1] = fin[e]
    comps = components(keep)
    report.append("forward stage B items: " + str([len(c) for c in comps]))
    nB = len(comps)
    for q, slots in enumerate(comps):
        src, fl = emit("fftf_B%d" % q, "const T* __restrict__ s, const ST st",
                       {s: keep[s] for s in slots},
                       lambda a: "s[%d * FFT_LS]" % a[1],
                       lambda slot, v: "st(%d, %s);" % (slot, v))
        out.append(src)
        report.append("  B%d: %d outputs, %d flops" % (q, len(slots), fl))
    out.append("#define FFTF_NA %d\n#define FFTF_NB %d\n" % (nA, nB))

    path = os.path.join(ROOT, "pyspeedy_b200/csrc/fft96_gen.cuh")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fp:
        fp.write("\n".join(out))
    print("\n".join(report))
    print("wrote", path)


if __name__ == "__main__":
    main()
