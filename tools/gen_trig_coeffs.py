#!/usr/bin/env python3
"""Prints the coefficient tables of the deterministic acos / cos used by FastEigen3x3's restatements
(oracle/oracle.c and cupoch_b200/csrc/cphb_eigen3.cuh carry the SAME tables, written out twice on purpose:
product and oracle share a specification, not code).  Exact rationals -> nearest double -> C99 hex floats."""
from fractions import Fraction
from math import factorial


def asin_coeffs(n):  # asin(z) = z * sum_k c_k z^(2k),  c_k = (2k)! / (4^k (k!)^2 (2k+1))
    return [Fraction(factorial(2 * k), 4 ** k * factorial(k) ** 2 * (2 * k + 1)) for k in range(n)]


def cos_coeffs(n):   # cos(u) = sum_k (-1)^k u^(2k) / (2k)!
    return [Fraction((-1) ** k, factorial(2 * k)) for k in range(n)]


def sin_coeffs(n):   # sin(u) = u * sum_k (-1)^k u^(2k) / (2k+1)!
    return [Fraction((-1) ** k, factorial(2 * k + 1)) for k in range(n)]


def table(name, cs):
    print("static const double %s[%d] = {" % (name, len(cs)))
    for c in cs:
        print("    %s," % float(c).hex())
    print("};")


if __name__ == "__main__":
    table("DT_ASIN", asin_coeffs(26))
    table("DT_COS", cos_coeffs(12))
    table("DT_SIN", sin_coeffs(12))
