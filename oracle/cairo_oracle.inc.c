/*
 * cairo_oracle.inc.c -- CPU ORACLE (test infrastructure, NOT the product), included by
 * stark252_oracle.c: restatement of the Cairo AIR pieces that sit between the LDE + commitment
 * calls of `prove`:
 *   build_auxiliary_trace        src/cairo/air.rs:660-729 (and :488-588)
 *   CairoAIR::compute_transition src/cairo/air.rs:743-767, 869-1160
 *   boundary_constraints         src/cairo/air.rs:777-849
 *   ConstraintEvaluator::evaluate src/starks/constraints/evaluator.rs:40-262
 * PARITY: pinned on benches/proofs/fibonacci_70000.proof -- with these functions the oracle
 * reproduces lde_trace_merkle_roots[1] (auxiliary trace) and composition_poly_root of that file
 * (tests/test_cairo_golden.py; the 2^19-row run is tools/cairo_golden_check.py).
 */

enum {
    C_F_DST_FP = 0, C_F_OP_0_FP = 1, C_F_OP_1_VAL = 2, C_F_OP_1_FP = 3, C_F_OP_1_AP = 4, C_F_RES_ADD = 5, C_F_RES_MUL = 6,
    C_F_PC_ABS = 7, C_F_PC_REL = 8, C_F_PC_JNZ = 9, C_F_AP_ADD = 10, C_F_AP_ONE = 11, C_F_OPC_CALL = 12, C_F_OPC_RET = 13,
    C_F_OPC_AEQ = 14,
    C_RES = 16, C_AP = 17, C_FP = 18, C_PC = 19, C_DST_ADDR = 20, C_OP0_ADDR = 21, C_OP1_ADDR = 22, C_INST = 23, C_DST = 24,
    C_OP0 = 25, C_OP1 = 26, C_OFF_DST = 27, C_OFF_OP0 = 28, C_OFF_OP1 = 29, C_T0 = 30, C_T1 = 31, C_MUL = 32, C_SELECTOR = 33,
    C_RC_0 = 34, C_RC_VALUE = 42,
    /* auxiliary columns, indices for the layout WITH the range-check builtin (air.rs:127-151) */
    C_RANGE_CHECK_COL_1 = 43, C_MEMORY_ADDR_SORTED_0 = 46, C_MEMORY_VALUES_SORTED_0 = 50, C_PERMUTATION_ARGUMENT_COL_0 = 54,
    C_PERMUTATION_ARGUMENT_RANGE_CHECK_COL_1 = 58,
    C_BUILTIN_OFFSET = 9, C_AUX_COLS = 18, C_N_TRANSITION = 49
};

typedef struct { uint64_t key; size_t idx; } sort_item;
static int sort_item_cmp(const void *a, const void *b) {
    const sort_item *x = a, *y = b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);     /* stable: Rust's sort_by is stable */
}
static inline uint64_t fe_low64(const fe *m) { fe c = fe_from_mont(m); return c.l[0]; }

static void batch_inverse(fe *v, size_t n) {   /* FieldElement::inplace_batch_inverse */
    if (!n) return;
    fe *pre = malloc(n * sizeof(fe));
    fe acc = ONE;
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; acc = fe_mul(&acc, &v[i]); }
    fe inv = fe_inv(&acc);
    for (size_t i = n; i-- > 0;) {
        fe t = fe_mul(&inv, &pre[i]);
        inv = fe_mul(&inv, &v[i]);
        v[i] = t;
    }
    free(pre);
}

/* build_auxiliary_trace (air.rs:660-729).  main: row-major n x n_cols; pub_addrs/pub_vals: the public
 * memory in address order (get_pub_memory_addrs, air.rs:508-527); rap = alpha_memory, z_memory,
 * z_range_check; aux_out: row-major n x 18. */
int o_cairo_build_aux_trace(const fe_lw *main, size_t n, size_t n_cols, const uint64_t *pub_addrs, const fe_lw *pub_vals,
                            size_t n_pub, const fe_lw *rap, fe_lw *aux_out) {
    if (!n || n_cols < 34 || 4 * n < n_pub) return -1;
    const fe alpha = lw_in(&rap[0]), z = lw_in(&rap[1]), zrc = lw_in(&rap[2]);
    const size_t L = 4 * n, R = 3 * n;
    static const int ACOL[4] = {C_PC, C_DST_ADDR, C_OP0_ADDR, C_OP1_ADDR}, VCOL[4] = {C_INST, C_DST, C_OP0, C_OP1};
    fe *a = malloc(L * sizeof(fe)), *v = malloc(L * sizeof(fe)), *as = malloc(L * sizeof(fe)), *vs = malloc(L * sizeof(fe));
    fe *den = malloc(L * sizeof(fe)), *perm = malloc(L * sizeof(fe));
    sort_item *items = malloc(L * sizeof(sort_item));
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 4; ++k) {
            a[4 * i + k] = lw_in(&main[i * n_cols + ACOL[k]]);
            v[4 * i + k] = lw_in(&main[i * n_cols + VCOL[k]]);
        }
    /* add_pub_memory_in_public_input_section (air.rs:488-506) */
    for (size_t i = 0; i < L; ++i) { as[i] = a[i]; vs[i] = v[i]; }
    for (size_t k = 0; k < n_pub; ++k) {
        as[L - n_pub + k] = fe_from_u64(pub_addrs[k]);
        vs[L - n_pub + k] = lw_in(&pub_vals[k]);
    }
    /* sort_columns_by_memory_address (air.rs:529-533): stable, by representative (addresses are words) */
    for (size_t i = 0; i < L; ++i) { items[i].key = fe_low64(&as[i]); items[i].idx = i; }
    qsort(items, L, sizeof(sort_item), sort_item_cmp);
    {
        fe *ta = malloc(L * sizeof(fe)), *tv = malloc(L * sizeof(fe));
        for (size_t i = 0; i < L; ++i) { ta[i] = as[items[i].idx]; tv[i] = vs[items[i].idx]; }
        memcpy(as, ta, L * sizeof(fe)); memcpy(vs, tv, L * sizeof(fe));
        free(ta); free(tv);
    }
    /* generate_memory_permutation_argument_column (air.rs:535-563) */
    for (size_t i = 0; i < L; ++i) { fe t = fe_mul(&alpha, &vs[i]); t = fe_add(&as[i], &t); den[i] = fe_sub(&z, &t); }
    batch_inverse(den, L);
    {
        fe prod = ONE;
        for (size_t i = 0; i < L; ++i) {
            fe t = fe_mul(&alpha, &v[i]); t = fe_add(&a[i], &t); t = fe_sub(&z, &t);
            t = fe_mul(&t, &den[i]);
            prod = fe_mul(&prod, &t);
            perm[i] = prod;
        }
    }
    /* range check (air.rs:683-700, 564-588) */
    fe *off = malloc(R * sizeof(fe)), *offs = malloc(R * sizeof(fe)), *rden = malloc(R * sizeof(fe)), *rperm = malloc(R * sizeof(fe));
    uint16_t *sorted = malloc(R * sizeof(uint16_t));
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) { off[3 * i + k] = lw_in(&main[i * n_cols + C_OFF_DST + k]); sorted[3 * i + k] = (uint16_t)fe_low64(&off[3 * i + k]); }
    {   /* counting sort of u16 values */
        size_t *cnt = calloc(65536, sizeof(size_t));
        for (size_t i = 0; i < R; ++i) cnt[sorted[i]]++;
        size_t p = 0;
        for (size_t val = 0; val < 65536; ++val) for (size_t k = 0; k < cnt[val]; ++k) sorted[p++] = (uint16_t)val;
        free(cnt);
    }
    for (size_t i = 0; i < R; ++i) { offs[i] = fe_from_u64(sorted[i]); rden[i] = fe_sub(&zrc, &offs[i]); }
    batch_inverse(rden, R);
    {
        fe prod = ONE;
        for (size_t i = 0; i < R; ++i) {
            fe t = fe_sub(&zrc, &off[i]);
            prod = fe_mul(&prod, &t);
            prod = fe_mul(&prod, &rden[i]);
            rperm[i] = prod;
        }
    }
    for (size_t i = 0; i < n; ++i) {
        fe_lw *r = aux_out + i * C_AUX_COLS;
        for (int k = 0; k < 3; ++k) lw_out(&offs[3 * i + k], &r[k]);
        for (int k = 0; k < 4; ++k) lw_out(&as[4 * i + k], &r[3 + k]);
        for (int k = 0; k < 4; ++k) lw_out(&vs[4 * i + k], &r[7 + k]);
        for (int k = 0; k < 4; ++k) lw_out(&perm[4 * i + k], &r[11 + k]);
        for (int k = 0; k < 3; ++k) lw_out(&rperm[3 * i + k], &r[15 + k]);
    }
    free(a); free(v); free(as); free(vs); free(den); free(perm); free(items);
    free(off); free(offs); free(rden); free(rperm); free(sorted);
    return 0;
}

/* CairoAIR::compute_transition (air.rs:743-767).  cur/nxt: one frame row each (main then aux columns);
 * bo = builtin offset subtracted from the auxiliary column indices (9 without the range-check builtin). */
static inline fe F_add(fe a, fe b) { return fe_add(&a, &b); }
static inline fe F_sub(fe a, fe b) { return fe_sub(&a, &b); }
static inline fe F_mul(fe a, fe b) { return fe_mul(&a, &b); }
static void cairo_transition(const fe *cur, const fe *nxt, const fe *rap, int has_rc, fe *c) {
    const int bo = has_rc ? 0 : C_BUILTIN_OFFSET;
    const fe one = ONE, two = fe_from_u64(2);
    /* compute_instr_constraints (air.rs:869-898) */
    for (int i = 0; i < 15; ++i) c[i] = F_mul(cur[i], F_sub(cur[i], one));
    c[15] = cur[15];
    fe f0 = ZERO;
    for (int i = 14; i >= 0; --i) f0 = F_add(cur[i], F_mul(two, f0));
    const fe b16 = fe_from_u64(1ULL << 16), b32 = fe_from_u64(1ULL << 32), b48 = fe_from_u64(1ULL << 48), b15 = fe_from_u64(1ULL << 15);
    c[16] = F_sub(F_add(F_add(F_add(cur[C_OFF_DST], F_mul(b16, cur[C_OFF_OP0])), F_mul(b32, cur[C_OFF_OP1])), F_mul(b48, f0)), cur[C_INST]);
    /* compute_operand_constraints (air.rs:900-927) */
    const fe ap = cur[C_AP], fp = cur[C_FP], pc = cur[C_PC];
    c[17] = F_sub(F_add(F_add(F_mul(cur[C_F_DST_FP], fp), F_mul(F_sub(one, cur[C_F_DST_FP]), ap)), F_sub(cur[C_OFF_DST], b15)), cur[C_DST_ADDR]);
    c[18] = F_sub(F_add(F_add(F_mul(cur[C_F_OP_0_FP], fp), F_mul(F_sub(one, cur[C_F_OP_0_FP]), ap)), F_sub(cur[C_OFF_OP0], b15)), cur[C_OP0_ADDR]);
    {
        fe rest = F_sub(F_sub(F_sub(one, cur[C_F_OP_1_VAL]), cur[C_F_OP_1_AP]), cur[C_F_OP_1_FP]);
        fe s = F_add(F_add(F_mul(cur[C_F_OP_1_VAL], pc), F_mul(cur[C_F_OP_1_AP], ap)), F_mul(cur[C_F_OP_1_FP], fp));
        s = F_add(s, F_mul(rest, cur[C_OP0]));
        s = F_add(s, F_sub(cur[C_OFF_OP1], b15));
        c[19] = F_sub(s, cur[C_OP1_ADDR]);
    }
    /* compute_register_constraints (air.rs:929-964) */
    const fe inst_size = F_add(cur[C_F_OP_1_VAL], one);
    c[20] = F_sub(F_add(F_add(F_add(ap, F_mul(cur[C_F_AP_ADD], cur[C_RES])), cur[C_F_AP_ONE]), F_mul(cur[C_F_OPC_CALL], two)), nxt[C_AP]);
    c[21] = F_sub(F_add(F_add(F_mul(cur[C_F_OPC_RET], cur[C_DST]), F_mul(cur[C_F_OPC_CALL], F_add(ap, two))),
                        F_mul(F_sub(F_sub(one, cur[C_F_OPC_RET]), cur[C_F_OPC_CALL]), fp)), nxt[C_FP]);
    c[22] = F_mul(F_sub(cur[C_T1], cur[C_F_PC_JNZ]), F_sub(nxt[C_PC], F_add(pc, inst_size)));
    {
        fe lhs = F_add(F_mul(cur[C_T0], F_sub(nxt[C_PC], F_add(pc, cur[C_OP1]))), F_mul(F_sub(one, cur[C_F_PC_JNZ]), nxt[C_PC]));
        fe reg = F_sub(F_sub(F_sub(one, cur[C_F_PC_ABS]), cur[C_F_PC_REL]), cur[C_F_PC_JNZ]);
        fe rhs = F_add(F_add(F_mul(reg, F_add(pc, inst_size)), F_mul(cur[C_F_PC_ABS], cur[C_RES])), F_mul(cur[C_F_PC_REL], F_add(pc, cur[C_RES])));
        c[23] = F_sub(lhs, rhs);
    }
    c[24] = F_sub(F_mul(cur[C_F_PC_JNZ], cur[C_DST]), cur[C_T0]);
    c[25] = F_sub(F_mul(cur[C_T0], cur[C_RES]), cur[C_T1]);
    /* compute_opcode_constraints (air.rs:966-984) */
    c[26] = F_sub(cur[C_MUL], F_mul(cur[C_OP0], cur[C_OP1]));
    {
        fe rest = F_sub(F_sub(F_sub(one, cur[C_F_RES_ADD]), cur[C_F_RES_MUL]), cur[C_F_PC_JNZ]);
        fe s = F_add(F_add(F_mul(cur[C_F_RES_ADD], F_add(cur[C_OP0], cur[C_OP1])), F_mul(cur[C_F_RES_MUL], cur[C_MUL])), F_mul(rest, cur[C_OP1]));
        c[27] = F_sub(s, F_mul(F_sub(one, cur[C_F_PC_JNZ]), cur[C_RES]));
    }
    c[28] = F_mul(cur[C_F_OPC_CALL], F_sub(cur[C_DST], fp));
    c[29] = F_mul(cur[C_F_OPC_CALL], F_sub(cur[C_OP0], F_add(pc, inst_size)));
    c[30] = F_mul(cur[C_F_OPC_AEQ], F_sub(cur[C_DST], cur[C_RES]));
    /* enforce_selector (air.rs:986-991) */
    for (int i = 16; i <= 30; ++i) c[i] = F_mul(c[i], cur[C_SELECTOR]);
    /* memory_is_increasing (air.rs:993-1049) */
    const fe *A = cur + C_MEMORY_ADDR_SORTED_0 - bo, *V = cur + C_MEMORY_VALUES_SORTED_0 - bo;
    const fe a_next = nxt[C_MEMORY_ADDR_SORTED_0 - bo], v_next = nxt[C_MEMORY_VALUES_SORTED_0 - bo];
    for (int k = 0; k < 4; ++k) {
        const fe a0 = A[k], a1 = k < 3 ? A[k + 1] : a_next, v0 = V[k], v1 = k < 3 ? V[k + 1] : v_next;
        const fe step = F_sub(F_sub(a1, a0), one);
        c[31 + k] = F_mul(F_sub(a0, a1), step);
        c[35 + k] = F_mul(F_sub(v0, v1), step);
    }
    /* permutation_argument (air.rs:1051-1096) */
    {
        const fe z = rap[1], alpha = rap[0];
        const fe *Pm = cur + C_PERMUTATION_ARGUMENT_COL_0 - bo;
        const fe p_next = nxt[C_PERMUTATION_ARGUMENT_COL_0 - bo];
        const fe a_[4] = {nxt[C_PC], cur[C_DST_ADDR], cur[C_OP0_ADDR], cur[C_OP1_ADDR]};
        const fe v_[4] = {nxt[C_INST], cur[C_DST], cur[C_OP0], cur[C_OP1]};
        for (int k = 1; k <= 4; ++k) {   /* constraint k-1 links p_{k-1} -> p_k (p_4 = p0_next) */
            const fe apk = k < 4 ? A[k] : a_next, vpk = k < 4 ? V[k] : v_next, pk = k < 4 ? Pm[k] : p_next;
            const fe ak = a_[k & 3], vk = v_[k & 3];
            fe l = F_mul(F_sub(z, F_add(apk, F_mul(alpha, vpk))), pk);
            fe r = F_mul(F_sub(z, F_add(ak, F_mul(alpha, vk))), Pm[k - 1]);
            c[39 + k - 1] = F_sub(l, r);
        }
    }
    /* permutation_argument_range_check (air.rs:1098-1139) */
    {
        const fe z = rap[2];
        const fe *Rc = cur + C_RANGE_CHECK_COL_1 - bo, *Pr = cur + C_PERMUTATION_ARGUMENT_RANGE_CHECK_COL_1 - bo;
        const fe rc_next = nxt[C_RANGE_CHECK_COL_1 - bo], pr_next = nxt[C_PERMUTATION_ARGUMENT_RANGE_CHECK_COL_1 - bo];
        for (int k = 0; k < 3; ++k) {
            const fe a0 = Rc[k], a1 = k < 2 ? Rc[k + 1] : rc_next;
            c[43 + k] = F_mul(F_sub(a0, a1), F_sub(F_sub(a1, a0), one));
        }
        const fe a_[3] = {nxt[C_OFF_DST], cur[C_OFF_OP0], cur[C_OFF_OP1]};
        for (int k = 1; k <= 3; ++k) {
            const fe apk = k < 3 ? Rc[k] : rc_next, pk = k < 3 ? Pr[k] : pr_next, ak = a_[k % 3];
            c[46 + k - 1] = F_sub(F_mul(F_sub(z, apk), pk), F_mul(F_sub(z, ak), Pr[k - 1]));
        }
    }
    if (has_rc) {   /* range_check_builtin (air.rs:1141-1160) */
        fe acc = ZERO;
        for (int k = 7; k >= 0; --k) acc = F_add(cur[C_RC_0 + k], F_mul(b16, acc));
        c[49] = F_sub(acc, cur[C_RC_VALUE]);
    }
}

static const uint8_t CAIRO_DEGREES[50] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3,
                                          2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1};
static const uint8_t CAIRO_EXEMPTIONS[50] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0,
                                             0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0};

/* CairoAIR::compute_transition for one frame (test hook) */
void o_cairo_compute_transition(const fe_lw *cur, const fe_lw *nxt, size_t n_cols, const fe_lw *rap, int has_rc, fe_lw *out) {
    fe c0[70], c1[70], r[3], c[50];
    for (size_t j = 0; j < n_cols && j < 70; ++j) { c0[j] = lw_in(&cur[j]); c1[j] = lw_in(&nxt[j]); }
    for (int k = 0; k < 3; ++k) r[k] = lw_in(&rap[k]);
    cairo_transition(c0, c1, r, has_rc, c);
    for (int k = 0; k < C_N_TRANSITION + (has_rc ? 1 : 0); ++k) lw_out(&c[k], &out[k]);
}

typedef struct {
    const fe *lde; size_t n_cols, n, m, blowup; int has_rc;
    fe rap[3], offset, w;                 /* w = primitive m-th root */
    size_t nb; uint64_t bstep[8]; unsigned bcol[8]; fe bval[8], bcoef[8][2];
    fe tcoef[50][2];
    fe *out;
    size_t lo, hi;
} cairo_eval_job;

static void *cairo_eval_worker(void *arg) {
    cairo_eval_job *J = arg;
    const size_t n = J->n, m = J->m, b = J->blowup, nc = J->n_cols;
    const int nt = C_N_TRANSITION + (J->has_rc ? 1 : 0);
    const size_t cnt = J->hi - J->lo;
    /* trace_primitive_root g = w^blowup; boundary points g^step */
    fe g = fe_pow_u64(&J->w, b);
    fe bpoint[8];
    for (size_t k = 0; k < J->nb; ++k) bpoint[k] = fe_pow_u64(&g, J->bstep[k]);
    fe g_last = fe_pow_u64(&g, n - 1);          /* exemption polynomial x - g^(n-1) (traits.rs:42-76) */
    /* zerofier inverses 1/(d^n - 1): d^n = offset^n * (w^n)^i takes `blowup` values (evaluator.rs:150-163) */
    fe *zinv = malloc(b * sizeof(fe)), *dn = malloc(b * sizeof(fe));
    {
        fe hn = fe_pow_u64(&J->offset, n), wn = fe_pow_u64(&J->w, n), cur = hn;
        for (size_t r = 0; r < b; ++r) { dn[r] = cur; zinv[r] = fe_sub(&cur, &ONE); cur = fe_mul(&cur, &wn); }
        batch_inverse(zinv, b);
    }
    /* boundary zerofier inverses for this slice (evaluator.rs:58-72) */
    fe *binv = malloc(J->nb * cnt * sizeof(fe)), *dom = malloc(cnt * sizeof(fe));
    {
        fe d = fe_pow_u64(&J->w, J->lo); d = fe_mul(&d, &J->offset);
        for (size_t i = 0; i < cnt; ++i) { dom[i] = d; d = fe_mul(&d, &J->w); }
        for (size_t k = 0; k < J->nb; ++k)
            for (size_t i = 0; i < cnt; ++i) binv[k * cnt + i] = fe_sub(&dom[i], &bpoint[k]);
        batch_inverse(binv, J->nb * cnt);
    }
    fe cur[70], nxt[70], c[50];
    for (size_t ii = 0; ii < cnt; ++ii) {
        const size_t i = J->lo + ii, i2 = (i + b) % m;      /* Frame::read_from_trace (frame.rs:43-63), offsets [0, 1] */
        for (size_t j = 0; j < nc; ++j) { cur[j] = J->lde[j * m + i]; nxt[j] = J->lde[j * m + i2]; }
        const fe d = dom[ii];
        /* degree adjustments: bound = 2n; boundary d^(bound - n) = d^n; transition degree k: d^(2n - n(k-1)) */
        const fe dN = dn[i % b], d2N = fe_mul(&dN, &dN);
        fe acc = ZERO;
        for (size_t k = 0; k < J->nb; ++k) {
            fe coef = F_add(F_mul(J->bcoef[k][0], dN), J->bcoef[k][1]);
            fe t = F_mul(F_mul(binv[k * cnt + ii], coef), F_sub(cur[J->bcol[k]], J->bval[k]));
            acc = fe_add(&acc, &t);
        }
        cairo_transition(cur, nxt, J->rap, J->has_rc, c);
        const fe ex = fe_sub(&d, &g_last);
        for (int k = 0; k < nt; ++k) {
            const fe adj = CAIRO_DEGREES[k] == 1 ? d2N : CAIRO_DEGREES[k] == 2 ? dN : ONE;
            fe t = F_mul(F_mul(zinv[i % b], F_add(F_mul(J->tcoef[k][0], adj), J->tcoef[k][1])), c[k]);
            if (CAIRO_EXEMPTIONS[k]) t = fe_mul(&t, &ex);
            acc = fe_add(&acc, &t);
        }
        J->out[i] = acc;
    }
    free(zinv); free(dn); free(binv); free(dom);
    return NULL;
}

/* ConstraintEvaluator::evaluate for CairoAIR (evaluator.rs:40-262): lde = column-major n_cols x m (main
 * then auxiliary columns); boundary constraints as (col, step, value) triples in the order of
 * CairoAIR::boundary_constraints; *_coeffs[k] = (alpha_k, beta_k).  out: m evaluations of the
 * composition polynomial on the LDE coset. */
int o_cairo_constraint_evaluations(const fe_lw *lde, size_t n_cols, size_t n, size_t blowup, uint64_t coset_offset, int has_rc,
                                   const fe_lw *rap, size_t n_boundary, const uint64_t *bcols, const uint64_t *bsteps,
                                   const fe_lw *bvalues, const fe_lw *boundary_coeffs, const fe_lw *transition_coeffs,
                                   int threads, fe_lw *out) {
    const size_t m = n * blowup;
    if (!is_pow2(n) || !is_pow2(blowup) || n_boundary > 8 || n_cols > 70) return -1;
    fe *L = malloc(n_cols * m * sizeof(fe)), *res = malloc(m * sizeof(fe));
    if (!L || !res) { free(L); free(res); return -3; }
    for (size_t i = 0; i < n_cols * m; ++i) L[i] = lw_in(&lde[i]);
    cairo_eval_job base;
    memset(&base, 0, sizeof base);
    base.lde = L; base.n_cols = n_cols; base.n = n; base.m = m; base.blowup = blowup; base.has_rc = has_rc;
    for (int k = 0; k < 3; ++k) base.rap[k] = lw_in(&rap[k]);
    base.offset = fe_from_u64(coset_offset);
    primitive_root(ilog2(m), &base.w);
    base.nb = n_boundary;
    for (size_t k = 0; k < n_boundary; ++k) {
        base.bcol[k] = (unsigned)bcols[k]; base.bstep[k] = bsteps[k]; base.bval[k] = lw_in(&bvalues[k]);
        base.bcoef[k][0] = lw_in(&boundary_coeffs[2 * k]); base.bcoef[k][1] = lw_in(&boundary_coeffs[2 * k + 1]);
    }
    const int nt = C_N_TRANSITION + (has_rc ? 1 : 0);
    for (int k = 0; k < nt; ++k) { base.tcoef[k][0] = lw_in(&transition_coeffs[2 * k]); base.tcoef[k][1] = lw_in(&transition_coeffs[2 * k + 1]); }
    base.out = res;
    int nth = threads > 0 ? threads : 1;
    if (nth > 64) nth = 64;
    if ((size_t)nth > m) nth = (int)m;
    cairo_eval_job jobs[64];
    pthread_t tid[64];
    for (int t = 0; t < nth; ++t) {
        jobs[t] = base;
        jobs[t].lo = m * t / nth; jobs[t].hi = m * (t + 1) / nth;
        if (t) pthread_create(&tid[t], NULL, cairo_eval_worker, &jobs[t]);
    }
    cairo_eval_worker(&jobs[0]);
    for (int t = 1; t < nth; ++t) pthread_join(tid[t], NULL);
    for (size_t i = 0; i < m; ++i) lw_out(&res[i], &out[i]);
    free(L); free(res);
    return 0;
}
