"""Integer-pipe peak probes on the local GPU (roofline denominators missing from MEASURED_PEAKS.json)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N


def main():
    ctx = P.Context(0)
    arr = (C.c_double * 8)()
    ctx.check(N.lib().s252_microbench_int_pipes(ctx.handle, arr))
    g = C.c_double()
    ctx.check(N.lib().s252_microbench_fe_mul(ctx.handle, C.byref(g)))
    k = C.c_double()
    ctx.check(N.lib().s252_microbench_keccak(ctx.handle, C.byref(k)))
    out = {"imad_wide_gops": arr[0], "lop3_gops": arr[1], "shf_gops": arr[2], "iadd_carry_gops": arr[3],
           "imad_wide_plus_lop3_gops": arr[4], "imad_wide_carry_rows_gops": arr[5], "imad_lo_gops": arr[6],
           "imad_wide_same_parity_gops": arr[7], "fe_mul_gmuls": g.value, "keccak_gperms": k.value,
           "note": "every multiplicand is data-dependent; the round-1 probe multiplied loop-invariant registers, ptxas hoisted the "
                   "products and the loop measured IADD3 pairs (18.3 T = 2 adds per 'MAD' at 1 instr/clk/SMSP)"}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
