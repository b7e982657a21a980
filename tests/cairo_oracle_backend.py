"""CPU double of lambdaworks_cairo_prover_b200.cairo_distributed.GpuCairoBackend on the oracle (TEST INFRASTRUCTURE):
the same interface, CPU tensors holding LW elements, arithmetic from oracle/.  With it the orchestration of the
sharded Cairo prover (exchanges, halos, gathers, transcript replay, openings, serialization) runs under gloo
on CPU ranks and must produce the oracle prover's bytes.  Row-block steps are plain python-integer loops:
sizes are tiny."""
import numpy as np
import torch

from oracle import pyoracle as O
from oracle.cairo_prover import boundary_constraints
from oracle.cairo_verifier import DEGREES, EXEMPTIONS

P = O.P


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64))


def _a(t):
    return t.contiguous().numpy().view(np.uint64)


class _Handle:
    def __init__(self, coeffs, n_cols):
        self.coeffs, self.n_cols = coeffs, n_cols

    def free(self):
        pass

    def coefficients(self, j):
        return self.coeffs[j]


class OracleCairoBackend:
    device = torch.device("cpu")

    def __init__(self, trace_table, pub):
        self.table, self.pub = trace_table, pub          # [N, c, 4] LW: stands in for the pinned host table

    # ---- plumbing used by distributed.exchange_and_commit / ShardedCommit
    def sync(self):
        pass

    before_collective = after_collective = sync

    def commit_block(self, cols):
        arr = _a(cols)
        nodes, root = O.commit_columns(arr)
        return {"cols": arr, "nodes": nodes}, root

    def open_block(self, block, local_idx):
        return [block["cols"][:, i] for i in local_idx], [O.merkle_path(block["nodes"], i) for i in local_idx]

    @staticmethod
    def keccak(data):
        return O.keccak256(data)

    # ---- transcript
    def transcript(self):
        return O.Transcript()

    def to_field(self, t):
        return t.to_field()

    def to_usize(self, t):
        return t.to_usize()

    # ---- round 1
    def main_lde(self, trace, lo, hi, opts):
        sub = np.ascontiguousarray(self.table[:, lo:hi])
        r = O.interpolate_and_commit(sub, opts.blowup_factor, opts.coset_offset, want_nodes=False)
        return _Handle(r["coeffs"], hi - lo), _t(r["lde"]), _t(sub.transpose(1, 0, 2))

    def new_tensor(self, shape):
        return torch.zeros(shape, dtype=torch.int64)

    def aux_trace(self, trace, aux_in, rap):
        n = aux_in.shape[1]
        fake = np.zeros((n, 34, 4), dtype=np.uint64)
        fake[:, 19:30] = _a(aux_in).transpose(1, 0, 2)
        addrs = sorted(self.pub.public_memory)
        aux = O.cairo_build_aux_trace(fake, addrs, np.stack([self.pub.public_memory[a] for a in addrs]), rap)
        return _t(aux.transpose(1, 0, 2))

    def free_tensor(self, t):
        pass

    def cols_lde(self, cols, opts):
        sub = np.ascontiguousarray(_a(cols).transpose(1, 0, 2))
        r = O.interpolate_and_commit(sub, opts.blowup_factor, opts.coset_offset, want_nodes=False)
        return _Handle(r["coeffs"], cols.shape[0]), _t(r["lde"])

    def evaluate_at(self, handle, points):
        return np.stack([np.stack([O.poly_evaluate(handle.coefficients(j), p) for j in range(handle.n_cols)]) for p in points])

    # ---- round 2: ConstraintEvaluator::evaluate on a row block (evaluator.rs:40-262)
    def constraints_rows(self, trace, mblock, ablock, mhalo, ahalo, row0, rap, bco, tco, opts, out):
        b, h = opts.blowup_factor, opts.coset_offset
        rows = mblock.shape[1]
        n = trace.n_rows()
        m = n * b
        w = O.lw_to_int(O.primitive_root(m.bit_length() - 1))
        g = pow(w, b, P)
        cur = np.concatenate([_a(mblock), _a(ablock)])                       # [52, rows, 4]
        halo = np.concatenate([_a(mhalo), _a(ahalo)])                       # [52, b, 4]
        ext = np.concatenate([cur, halo], axis=1)
        bcs = boundary_constraints(self.pub, n, rap, False)
        ba, bb = [O.lw_to_int(x) for x in bco[:, 0]], [O.lw_to_int(x) for x in bco[:, 1]]
        ta, tb = [O.lw_to_int(x) for x in tco[:, 0]], [O.lw_to_int(x) for x in tco[:, 1]]
        res = np.zeros((rows, 4), dtype=np.uint64)
        for il in range(rows):
            i = row0 + il
            x = h * pow(w, i, P) % P
            xn = pow(x, n, P)
            zinv = pow((xn - 1) % P, -1, P)
            acc = 0
            for (col, step, value), a_, b_ in zip(bcs, ba, bb):
                acc += pow((x - pow(g, step, P)) % P, -1, P) * ((a_ * xn + b_) % P) * ((O.lw_to_int(cur[col, il]) - O.lw_to_int(value)) % P)
            c = O.lw_to_ints(O.cairo_compute_transition(np.ascontiguousarray(ext[:, il]), np.ascontiguousarray(ext[:, il + b]), rap))
            ex = (x - pow(g, n - 1, P)) % P
            for k, ck in enumerate(c):
                adj = {1: xn * xn % P, 2: xn, 3: 1}[DEGREES[k]]
                acc += zinv * ((ta[k] * adj + tb[k]) % P) * ck * (ex if EXEMPTIONS[k] else 1)
            res[il] = O.int_to_lw(acc % P)
        out.copy_(_t(res))

    def composition_commit(self, evals, n, opts, comp_lde_out):
        off = O.fe_from_u64(opts.coset_offset)
        hco = O.interpolate_offset_fft(_a(evals), off)
        assert not hco[2 * n:].any(), "composition polynomial exceeds its degree bound"
        h1, h2 = np.ascontiguousarray(hco[0:2 * n:2]), np.ascontiguousarray(hco[1:2 * n:2])
        lde = np.stack([O.evaluate_polynomial_on_lde_domain(h1, opts.blowup_factor, n, off),
                        O.evaluate_polynomial_on_lde_domain(h2, opts.blowup_factor, n, off)])
        nodes, root = O.commit_columns(lde)
        comp_lde_out.copy_(_t(lde))
        hnd = _Handle(np.stack([h1, h2]), 2)
        hnd.lde, hnd.nodes = lde, nodes
        return hnd, root

    def composition_lde(self, evals, n, opts):
        off = O.fe_from_u64(opts.coset_offset)
        hco = O.interpolate_offset_fft(_a(evals), off)
        assert not hco[2 * n:].any(), "composition polynomial exceeds its degree bound"
        h1, h2 = np.ascontiguousarray(hco[0:2 * n:2]), np.ascontiguousarray(hco[1:2 * n:2])
        lde = np.stack([O.evaluate_polynomial_on_lde_domain(h1, opts.blowup_factor, n, off),
                        O.evaluate_polynomial_on_lde_domain(h2, opts.blowup_factor, n, off)])
        return _Handle(np.stack([h1, h2]), 2), _t(lde)

    # ---- FRI building blocks (fri_distributed.py)
    def fri_fold_rows(self, v, s, i0, size, domain_size, k, zeta, coset_offset, out):
        w = O.lw_to_int(O.primitive_root(size.bit_length() - 1))
        hk = pow(coset_offset, 1 << k, P)
        z = O.lw_to_int(zeta)
        inv2 = pow(2, -1, P)
        va, sa = _a(v), _a(s)
        res = np.zeros((va.shape[0], 4), dtype=np.uint64)
        for j in range(va.shape[0]):
            a, b = O.lw_to_int(va[j]), O.lw_to_int(sa[j])
            x = hk * pow(w, i0 + j, P) % P
            res[j] = O.int_to_lw(((a + b) * inv2 + z * (a - b) * pow(2 * x, -1, P)) % P)
        out.copy_(_t(res))

    def fri_continue(self, evals, n_layers, t, coset_offset, k, domain_size):
        size = evals.shape[0]
        off = O.int_to_lw(pow(coset_offset, 1 << k, P))
        coeffs = O.interpolate_offset_fft(_a(evals), off)
        nz = np.nonzero(coeffs.any(axis=1))[0]
        coeffs = coeffs[: (nz[-1] + 1 if nz.size else 0)]
        last, roots, ev, nodes = O.fri_commit_phase(n_layers, coeffs, t, off, size)
        return (ev, nodes), last, roots

    def release_fri(self, fri):
        pass

    def grind_round(self, challenge, factor, base, part, parts, window_log):
        batch = 1 << 18
        for b in range(part, 1 << (window_log - 18), parts):
            for nonce in range(base + b * batch, base + (b + 1) * batch):
                if O.grinding_zeros(challenge, nonce) >= factor:
                    return nonce
        return (1 << 64) - 1

    @staticmethod
    def to_bytes_be(v):
        return O.fe_to_bytes_be(v)

    # ---- round 4: the DEEP polynomial on a row block (verifier.rs:526-557)
    def deep_rows(self, mblock, ablock, comp_lde, row0, n, z, ood, hz, gamma, gamma_p, tg, opts, out):
        b, h = opts.blowup_factor, opts.coset_offset
        rows, m = mblock.shape[1], comp_lde.shape[1]
        w = O.lw_to_int(O.primitive_root(m.bit_length() - 1))
        g = pow(w, b, P)
        cols = np.concatenate([_a(mblock), _a(ablock)])
        comp = _a(comp_lde)
        oodi = [[O.lw_to_int(v) for v in row] for row in ood]
        tgi = [O.lw_to_int(v) for v in tg]
        ga, gp, h1z, h2z = O.lw_to_int(gamma), O.lw_to_int(gamma_p), O.lw_to_int(hz[0]), O.lw_to_int(hz[1])
        res = np.zeros((rows, 4), dtype=np.uint64)
        for il in range(rows):
            i = row0 + il
            x = h * pow(w, i, P) % P
            dinv = [pow((x - z * pow(g, k, P)) % P, -1, P) for k in range(2)]
            acc = 0
            for j in range(cols.shape[0]):
                v = O.lw_to_int(cols[j, il])
                for k in range(2):
                    acc += (v - oodi[k][j]) * dinv[k] * tgi[2 * j + k]
            zinv = pow((x - z * z) % P, -1, P)
            acc += (O.lw_to_int(comp[0, i]) - h1z) * zinv * ga + (O.lw_to_int(comp[1, i]) - h2z) * zinv * gp
            res[il] = O.int_to_lw(acc % P)
        out.copy_(_t(res))

    def fri_commit_phase_evals(self, p0, layers, t, opts):
        off = O.fe_from_u64(opts.coset_offset)
        coeffs = O.interpolate_offset_fft(_a(p0), off)
        nz = np.nonzero(coeffs.any(axis=1))[0]
        coeffs = coeffs[: (nz[-1] + 1 if nz.size else 0)]
        last, roots, evals, nodes = O.fri_commit_phase(layers, coeffs, t, off, p0.shape[0])
        return (evals, nodes), last, roots

    def grind(self, challenge, factor):
        return O.generate_nonce_with_grinding(challenge, factor)

    def fri_query(self, fri, idx, layers, depth):
        evals, nodes = fri
        q = len(idx)
        ev = np.zeros((q, layers, 4), dtype=np.uint64)
        evs = np.zeros_like(ev)
        pa = np.zeros((q, layers, depth, 32), dtype=np.uint8)
        pas = np.zeros_like(pa)
        for a, iota in enumerate(idx):
            for k in range(layers):
                size = evals[k].shape[0]
                i, isym = iota % size, (iota + size // 2) % size
                ev[a, k], evs[a, k] = evals[k][i], evals[k][isym]
                pa[a, k, :depth - k] = O.merkle_path(nodes[k], i)
                pas[a, k, :depth - k] = O.merkle_path(nodes[k], isym)
        return ev, evs, pa, pas

    def commit_open(self, comp, idx, depth):
        rows = np.stack([comp.lde[:, i] for i in idx])
        paths = np.stack([O.merkle_path(comp.nodes, i) for i in idx])
        return rows, paths

    def release(self, fri, comp):
        pass
