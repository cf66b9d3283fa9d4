"""Coefficients of the degree-11 polynomial for exp(r) on |r| <= ln2/2 used by csrc/fastmath.cuh (tanh_bf), and its
error: interpolation through the Chebyshev nodes of the (slightly widened) interval in 60-digit arithmetic, the two top
coefficients rounded to doubles whose low word is zero (immediate operands).  Prints the hex literals and the maximum
relative error of the polynomial with the ROUNDED coefficients (evaluated exactly): 1.4e-17.

    python tests/prototypes/exp_poly.py
"""
import struct

import mpmath as mp

mp.mp.dps = 60
a = mp.log(2) / 2 * mp.mpf("1.001")
deg = 11
n = deg + 1
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = mp.e ** x
c = mp.lu_solve(A, b)


def high_word_only(v):
    bits = struct.unpack("<Q", struct.pack("<d", float(v)))[0]
    return struct.unpack("<d", struct.pack("<Q", (bits + 0x80000000) & 0xFFFFFFFF00000000))[0]


cd = [float(ci) for ci in c]
cd[10], cd[11] = high_word_only(c[10]), high_word_only(c[11])
for j, v in enumerate(cd):
    print(f"c{j:<2d} = {v.hex()}  ({v!r})")
worst = mp.mpf(0)
for i in range(4001):
    x = -a + 2 * a * i / 4000
    p = sum(mp.mpf(cd[j]) * x ** j for j in range(n))
    worst = max(worst, abs(p / mp.e ** x - 1))
print("max relative error:", mp.nstr(worst, 5))
