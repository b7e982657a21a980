"""The roofline denominators bench.py measures live (s252_microbench_int_pipes): guards against the probe being optimised away
again.  Round 1 multiplied loop-invariant registers, ptxas hoisted the products and the "IMAD.WIDE peak" was the rate of two adds
(twice the real one).  The assertions are ratios, so they hold at any clock."""
import ctypes as C

import pytest

import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N

pytestmark = pytest.mark.gpu


def test_integer_pipe_probes_measure_what_they_name():
    ctx = P.Context(0)
    arr = (C.c_double * 8)()
    ctx.check(N.lib().s252_microbench_int_pipes(ctx.handle, arr))
    imad_wide, lop3, shf, _, wide_plus_lop3, carry_rows, imad_lo, _ = list(arr)
    assert lop3 > 0 and abs(shf / lop3 - 1) < 0.1 and abs(imad_lo / lop3 - 1) < 0.1      # 64 lanes/clk/SM each
    assert 0.42 < imad_wide / lop3 < 0.58                                                    # IMAD.WIDE.U32: 32 lanes/clk/SM
    assert carry_rows < 1.05 * imad_wide                                                     # multiply-adds with carries are no faster
    assert wide_plus_lop3 > 1.7 * imad_wide                                                  # the logic op issues in the multiplier's shadow
    g = C.c_double()
    ctx.check(N.lib().s252_microbench_fe_mul(ctx.handle, C.byref(g)))
    # 80 algorithmic multiply-adds per field multiplication (72 issued as wide multiplies): the dependent stream runs near the roof
    assert 0.75 < g.value * 80 / imad_wide < 1.15
    ctx.close()
