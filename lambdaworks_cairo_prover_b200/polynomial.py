"""FFTPoly methods of lambdaworks-math's Polynomial, on the GPU.

Call sites in the reference: Polynomial::interpolate_fft (src/starks/trace.rs:107),
evaluate_offset_fft (src/starks/prover.rs:117, src/starks/fri/fri_commitment.rs:36),
interpolate_offset_fft (src/starks/constraints/evaluation_table.rs:32).
"""
import numpy as np

from . import _native as N
from . import felt


def _trimmed_len(coeffs):
    n = coeffs.shape[0]
    nz = np.flatnonzero(coeffs.any(axis=1))
    return int(nz[-1]) + 1 if nz.size else 0 if n else 0


class Polynomial:
    """Polynomial<FieldElement<Stark252PrimeField>>: `coefficients` low -> high, trailing zeros
    trimmed by the constructor (Polynomial::new)."""

    def __init__(self, coefficients):
        c = N.fe_array(np.asarray(coefficients, dtype=np.uint64).reshape(-1, 4))
        self.coefficients = np.ascontiguousarray(c[:_trimmed_len(c)])

    def coeff_len(self):
        return self.coefficients.shape[0]

    @staticmethod
    def interpolate_fft(fft_evals, ctx=None):
        ctx = ctx or N.default_context()
        ev = N.fe_array(np.asarray(fft_evals, dtype=np.uint64).reshape(-1, 4))
        out = np.empty_like(ev)
        ctx.check(N.lib().s252_interpolate_fft(ctx.handle, N.ptr(ev), ev.shape[0], N.ptr(out), N.HOST), N.FFTError)
        return Polynomial(out)

    @staticmethod
    def interpolate_offset_fft(fft_evals, offset, ctx=None):
        ctx = ctx or N.default_context()
        ev = N.fe_array(np.asarray(fft_evals, dtype=np.uint64).reshape(-1, 4))
        off = N.fe_array(offset)
        out = np.empty_like(ev)
        ctx.check(N.lib().s252_interpolate_offset_fft(ctx.handle, N.ptr(ev), ev.shape[0], N.ptr(off), N.ptr(out),
                                                      N.HOST), N.FFTError)
        return Polynomial(out)

    def evaluate_offset_fft(self, blowup_factor, domain_size, offset, ctx=None):
        """out[i] = p(offset * w_len^i), len = max(coeff_len, domain_size).next_power_of_two() * blowup."""
        ctx = ctx or N.default_context()
        off = N.fe_array(offset)
        n = self.coeff_len()
        length = N.lib().s252_evaluate_offset_fft_len(n, blowup_factor, domain_size or 0)
        out = np.empty((length, 4), dtype=np.uint64)
        ctx.check(N.lib().s252_evaluate_offset_fft(ctx.handle, N.ptr(self.coefficients) if n else None, n, blowup_factor,
                                                   domain_size or 0, N.ptr(off), N.ptr(out), length, N.HOST), N.FFTError)
        return out


def evaluate_polynomial_on_lde_domain(p, blowup_factor, domain_size, offset, ctx=None):
    """src/starks/prover.rs:106-123"""
    ctx = ctx or N.default_context()
    off = N.fe_array(offset)
    n = p.coeff_len()
    out = np.empty((domain_size * blowup_factor, 4), dtype=np.uint64)
    ctx.check(N.lib().s252_evaluate_polynomial_on_lde_domain(ctx.handle, N.ptr(p.coefficients) if n else None, n,
                                                             blowup_factor, domain_size, N.ptr(off), N.ptr(out), N.HOST),
              N.FFTError)
    return out
