"""The invariant the bulk / panel kernels' fast path rests on (floydwarshall_b200/csrc/fw_bulk.cuh):

    d = fma_rd(a, b, -o)          one DFMA, round toward minus infinity
    sign bit of d set  ==>  the reference's test  o < RN(a*b)  (Algorithms.hs:55,61) is false

checked here in exact rational arithmetic on generated binary64 values: the sign bit of RD(a*b - o) is set
exactly when the exact product is <= o (an exact tie gives -0 under round-down), and RN is monotone, so
RN(a*b) <= o.  The converse does not hold (exact a*b slightly above o can still round to o): those are
the filter's false candidates, which only cost a replay of the exact path."""
import math
from fractions import Fraction

from hypothesis import given, settings, strategies as st

finite_pos = st.floats(min_value=0.0, allow_nan=False, allow_infinity=False, allow_subnormal=True)


def sign_of_fma_rd(a: float, b: float, o: float) -> bool:
    """Sign bit of fma(a, b, -o) under round-toward-minus-infinity, for finite inputs."""
    exact = Fraction(a) * Fraction(b) - Fraction(o)
    if exact != 0:
        return exact < 0            # rounding never crosses zero
    # an exact zero result of a sum is -0 under round-down unless both addends are +0
    prod_is_pos_zero = (a == 0.0 or b == 0.0) and (math.copysign(1.0, a) * math.copysign(1.0, b) > 0)
    addend_is_pos_zero = (o == 0.0 and math.copysign(1.0, -o) > 0)
    return not (prod_is_pos_zero and addend_is_pos_zero)


@settings(max_examples=3000, deadline=None)
@given(finite_pos, finite_pos, finite_pos)
def test_set_sign_bit_proves_no_replacement(a, b, o):
    if sign_of_fma_rd(a, b, o):
        assert not (o < a * b)       # a * b in Python is the one rounded binary64 multiply of the reference


@settings(max_examples=2000, deadline=None)
@given(st.floats(min_value=1e-3, max_value=1e3), st.floats(min_value=1e-3, max_value=1e3),
       st.integers(min_value=-3, max_value=3))
def test_near_ties(a, b, ulps):
    """o within a few ulps of the rounded product: where the filter and the exact test can disagree."""
    p = a * b
    o = p
    for _ in range(abs(ulps)):
        o = math.nextafter(o, math.inf if ulps > 0 else 0.0)
    if sign_of_fma_rd(a, b, o):
        assert not (o < p)
    if o < p:                         # every real replacement is a candidate
        assert not sign_of_fma_rd(a, b, o)
