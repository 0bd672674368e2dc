#!/usr/bin/env python
"""Code generator for the in-register DFT butterflies of the two-pass FFT engine (fft_radix_gen.cuh).

For every radix R in RADICES it emits

    template <int S> HD void dft_R(const cplx* x, cplx* X);     // X[k] = sum_n x[n] exp(S*2*pi*i*n*k/R)

as straight-line code on scalar doubles (Cooley-Tukey R = A*B recursion down to hand-written 2/3/4/5
point butterflies).  Twiddle factors inside a butterfly are literals; multiplications by 1, -1, +-i and
(+-1+-i)/sqrt(2) are strength-reduced.  S (= -1 forward, +1 backward) is a template parameter that only
flips the sign of imaginary twiddle parts, so both directions come from the same text.
Run:  python gen_radix.py > fft_radix_gen.cuh
"""
import math
import sys

RADICES = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25]


class Gen:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def emit(self, s):
        self.lines.append("  " + s)

    # a complex value is a pair of C expressions (re, im) naming double variables
    def new(self, re_expr, im_expr):
        a = self.tmp()
        self.emit(f"const double {a}r = {re_expr}, {a}i = {im_expr};")
        return (a + "r", a + "i")

    def add(self, a, b):
        return self.new(f"{a[0]} + {b[0]}", f"{a[1]} + {b[1]}")

    def sub(self, a, b):
        return self.new(f"{a[0]} - {b[0]}", f"{a[1]} - {b[1]}")

    def muli(self, a):
        """a * (s*i)"""
        return self.new(f"-(s * {a[1]})", f"s * {a[0]}")

    def mul_tw(self, a, m, R):
        """a * exp(S*2*pi*i*m/R) with literal constants."""
        m %= R
        if m == 0:
            return a
        if (4 * m) % R == 0:
            q = (4 * m) // R            # quarter turns: 1 -> s*i, 2 -> -1, 3 -> -s*i
            if q == 1:
                return self.muli(a)
            if q == 2:
                return self.new(f"-{a[0]}", f"-{a[1]}")
            return self.new(f"s * {a[1]}", f"-(s * {a[0]})")
        c, sn = math.cos(2 * math.pi * m / R), math.sin(2 * math.pi * m / R)
        if (8 * m) % R == 0:
            # odd eighth turns: (+-1 +- s*i)/sqrt(2)
            h = repr(math.sqrt(0.5))
            sc = "+" if c > 0 else "-"
            ss = "" if sn > 0 else "-"
            # (c + i s sn)(ar + i ai) = c ar - s sn ai + i (c ai + s sn ar), |c| = |sn| = h
            return self.new(f"{h} * ({sc}{a[0]} - ({ss}s) * {a[1]})", f"{h} * ({sc}{a[1]} + ({ss}s) * {a[0]})")
        cs, ss = repr(c), repr(sn)
        return self.new(f"{cs} * {a[0]} - (s * {ss}) * {a[1]}", f"{cs} * {a[1]} + (s * {ss}) * {a[0]}")

    def dft(self, xs):
        R = len(xs)
        if R == 1:
            return list(xs)
        if R == 2:
            return [self.add(xs[0], xs[1]), self.sub(xs[0], xs[1])]
        if R == 4:
            t0, t1 = self.add(xs[0], xs[2]), self.sub(xs[0], xs[2])
            t2, t3 = self.add(xs[1], xs[3]), self.muli(self.sub(xs[1], xs[3]))
            return [self.add(t0, t2), self.add(t1, t3), self.sub(t0, t2), self.sub(t1, t3)]
        if R == 3:
            h = repr(math.sqrt(3) / 2)
            t1 = self.add(xs[1], xs[2])
            m = self.new(f"{xs[0][0]} - 0.5 * {t1[0]}", f"{xs[0][1]} - 0.5 * {t1[1]}")
            d0 = self.sub(xs[1], xs[2])
            d = self.new(f"-(s * {h}) * {d0[1]}", f"(s * {h}) * {d0[0]}")
            return [self.add(xs[0], t1), self.add(m, d), self.sub(m, d)]
        if R == 5:
            c1, c2 = repr(math.cos(2 * math.pi / 5)), repr(math.cos(4 * math.pi / 5))
            s1, s2 = repr(math.sin(2 * math.pi / 5)), repr(math.sin(4 * math.pi / 5))
            t1, t2 = self.add(xs[1], xs[4]), self.add(xs[2], xs[3])
            t3, t4 = self.sub(xs[1], xs[4]), self.sub(xs[2], xs[3])
            x0 = xs[0]
            m1 = self.new(f"{x0[0]} + {c1} * {t1[0]} + {c2} * {t2[0]}", f"{x0[1]} + {c1} * {t1[1]} + {c2} * {t2[1]}")
            m2 = self.new(f"{x0[0]} + {c2} * {t1[0]} + {c1} * {t2[0]}", f"{x0[1]} + {c2} * {t1[1]} + {c1} * {t2[1]}")
            n1 = self.new(f"-s * ({s1} * {t3[1]} + {s2} * {t4[1]})", f"s * ({s1} * {t3[0]} + {s2} * {t4[0]})")
            n2 = self.new(f"-s * ({s2} * {t3[1]} - {s1} * {t4[1]})", f"s * ({s2} * {t3[0]} - {s1} * {t4[0]})")
            X0 = self.new(f"{x0[0]} + {t1[0]} + {t2[0]}", f"{x0[1]} + {t1[1]} + {t2[1]}")
            return [X0, self.add(m1, n1), self.add(m2, n2), self.sub(m2, n2), self.sub(m1, n1)]
        # composite: R = A * B, x[B a + b];  X[c + A d] = sum_b W_R^{bc} W_B^{bd} (sum_a x[Ba+b] W_A^{ac})
        A = 4 if R % 4 == 0 else (2 if R % 2 == 0 else (3 if R % 3 == 0 else 5))
        if R % A != 0:
            raise ValueError(f"unsupported radix {R}")
        B = R // A
        y = [self.dft([xs[B * a + b] for a in range(A)]) for b in range(B)]      # y[b][c]
        X = [None] * R
        for c in range(A):
            col = self.dft([self.mul_tw(y[b][c], b * c, R) for b in range(B)])
            for d in range(B):
                X[c + A * d] = col[d]
        return X


def main():
    out = ["// GENERATED by gen_radix.py -- do not edit.  In-register DFT butterflies (straight-line code).",
           "#pragma once", '#include "fft_core.cuh"', "namespace dftk {", ""]
    for R in RADICES:
        g = Gen()
        xs = []
        for n in range(R):
            xs.append((f"x[{n}].x", f"x[{n}].y"))
        X = g.dft(xs)
        out.append(f"template <int S> HD void dft_{R}(const cplx* __restrict__ x, cplx* __restrict__ X) {{")
        out.append("  const double s = (double)S; (void)s;")
        out += g.lines
        for k in range(R):
            out.append(f"  X[{k}] = make_double2({X[k][0]}, {X[k][1]});")
        out.append("}")
        out.append("")
    out.append("template <int R, int S> HD void dft_r(const cplx* __restrict__ x, cplx* __restrict__ X) {")
    for R in RADICES:
        out.append(f"  if (R == {R}) dft_{R}<S>(x, X);")
    out.append("}")
    out.append("}  // namespace dftk")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
