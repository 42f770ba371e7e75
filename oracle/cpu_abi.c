/*
 * cpu_abi.c — CPU ORACLE build of the C ABI.  TEST INFRASTRUCTURE ONLY.
 *
 * Exports the entry points of include/ministark_b200.h on top of the CPU oracle (gl_oracle.c, included below as one
 * translation unit), so that the C++ host layer (include/ministark_prover.hpp, include/ministark_gpu.hpp) can be
 * linked, run and byte-compared WITHOUT a GPU:
 *   - tests/test_cpp_cpu_abi.py: the C++ prover's proof bytes == oracle/stark_oracle.cpu_prove's (the restated
 *     reference prover) in the CPU suite — the C++ host logic is then covered before any GPU run;
 *   - bench.py --impl reference: a prover compiled end to end as a second CPU baseline (`full_prove_compiled`).
 * It is a different library (oracle/libms_cpu_abi.so); the product (ministark_b200/, libministark_b200.so) never
 * loads it and has no CPU fallback.  "Device" pointers are plain host pointers here.
 *
 * Not provided (no CPU meaning): ms_ipc_*, ms_lde_batch_scatter, ms_eval_jit_check, ms_debug_lazy_ops.
 *
 * The fused evaluator (ms_eval_constraints*) is restated the way the reference's CPU evaluator works
 * (src/eval_cpu.rs:33-150): the flattened program is interpreted chunk by chunk, every instruction over a whole
 * chunk of domain points, divisions by batch inversion per chunk (eval_cpu.rs:280-295).
 */
#include "gl_oracle.c"

#include <stdarg.h>
#include <stdio.h>

#include "../include/ministark_b200.h"

struct ms_ctx {
    char err[512];
    uint64_t calls;
};

static int fail(ms_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}
static int bad_field(int f) { return f != MS_FIELD_FP && f != MS_FIELD_FQ3; }

/* MS_CPU_ABI_TRACE=1: one stderr line per heavy call (name, seconds) — the phase breakdown of a CPU prove */
#include <time.h>
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int done(ms_ctx *c, const char *what, double t0) {
    static int trace = -1;
    if (trace < 0) { const char *e = getenv("MS_CPU_ABI_TRACE"); trace = e && *e && *e != '0'; }
    if (trace) fprintf(stderr, "[cpu_abi] %-28s %9.4f s\n", what, now_s() - t0);
    c->calls++;
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ context / memory */
int ms_ctx_create(int device, ms_ctx **out) {
    (void)device;
    if (!out) return MS_ERR_INVALID;
    ms_ctx *c = (ms_ctx *)calloc(1, sizeof *c);
    if (!c) return MS_ERR_NOMEM;
    *out = c;
    return MS_OK;
}
int ms_ctx_destroy(ms_ctx *c) { free(c); return MS_OK; }
int ms_ctx_set_stream(ms_ctx *c, void *s) { (void)s; return c ? MS_OK : MS_ERR_INVALID; }
int ms_ctx_sync(ms_ctx *c) { return c ? MS_OK : MS_ERR_INVALID; }
const char *ms_last_error(ms_ctx *c) { return c ? c->err : "no context"; }
const char *ms_version(void) { return "ministark_b200 CPU oracle ABI (test infrastructure)"; }
int ms_set_option(ms_ctx *c, const char *name, int64_t v) { (void)name; (void)v; return c ? MS_OK : MS_ERR_INVALID; }
uint64_t ms_launch_count(ms_ctx *c) { (void)c; return 0; }   /* no kernels: a caller counting launches sees 0 */

int ms_alloc_device(ms_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return MS_ERR_INVALID;
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64)) return fail(c, MS_ERR_NOMEM, "out of host memory (%zu bytes)", bytes);
    *out = p;
    return MS_OK;
}
int ms_alloc_host_pinned(ms_ctx *c, size_t bytes, void **out) { return ms_alloc_device(c, bytes, out); }
int ms_free(ms_ctx *c, void *p) { (void)c; free(p); return MS_OK; }
int ms_copy(ms_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (bytes && (!dst || !src))) return MS_ERR_INVALID;
    memmove(dst, src, bytes);
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ transforms */
static int check_offset(ms_ctx *c, u64 off) {
    return (off == 0 || off >= GL_P) ? fail(c, MS_ERR_INVALID, "offset must be a non-zero canonical word") : MS_OK;
}
int ms_ntt_batch_to(ms_ctx *c, int field, const void *src, size_t ss, void *dst, size_t ds, unsigned ncols, unsigned log_n,
                    int direction, uint64_t offset) {
    if (!c || !src || !dst) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (log_n > 32) return fail(c, MS_ERR_INVALID, "log_n > 32");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    const size_t n = (size_t)1 << log_n;
    if (ncols > 1 && (ss < n || ds < n)) return fail(c, MS_ERR_INVALID, "column stride < 2^log_n");
    if (src == dst && ss != ds && ncols > 1) return fail(c, MS_ERR_INVALID, "in place with two different strides");
    const double t0 = now_s();
    if (src != dst)
        for (unsigned k = 0; k < ncols; k++)
            memmove((u64 *)dst + (size_t)k * ds * field, (const u64 *)src + (size_t)k * ss * field, n * field * 8);
    orc_ntt_columns((u64 *)dst, ds * field, ncols, (unsigned)field, log_n, offset, direction == MS_NTT_INVERSE);
    return done(c, direction == MS_NTT_INVERSE ? "ms_ntt_batch (inverse)" : "ms_ntt_batch (forward)", t0);
}
int ms_ntt_batch(ms_ctx *c, int field, void *data, size_t stride, unsigned ncols, unsigned log_n, int direction, uint64_t offset) {
    return ms_ntt_batch_to(c, field, data, stride, data, stride, ncols, log_n, direction, offset);
}
int ms_lde_batch(ms_ctx *c, int field, const void *coeffs, size_t is, void *evals, size_t os, unsigned ncols, unsigned log_n,
                 unsigned log_blowup, uint64_t offset, int bitrev_out) {
    if (!c || !coeffs || !evals) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (log_n + log_blowup > 32) return fail(c, MS_ERR_INVALID, "log_n + log_blowup > 32");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    if (coeffs == evals) return fail(c, MS_ERR_INVALID, "ms_lde_batch is out of place");
    const double t0 = now_s();
    orc_lde_columns((const u64 *)coeffs, is * field, (u64 *)evals, os * field, ncols, (unsigned)field, log_n, log_blowup, offset, bitrev_out);
    return done(c, "ms_lde_batch", t0);
}
int ms_bit_reverse(ms_ctx *c, int field, void *data, size_t stride, unsigned ncols, unsigned log_n) {
    if (!c || !data) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    for (unsigned k = 0; k < ncols; k++) orc_bit_reverse((u64 *)data + (size_t)k * stride * field, (unsigned)field, log_n);
    return MS_OK;
}

/* GpuFft / GpuIfft: encode columns, execute all */
struct ms_ntt_plan {
    ms_ctx *ctx;
    int field, direction;
    unsigned log_n;
    u64 offset;
    void **cols;
    size_t ncols, cap;
};
int ms_ntt_plan_create(ms_ctx *c, int field, unsigned log_n, int direction, uint64_t offset, ms_ntt_plan **out) {
    if (!c || !out) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (log_n > 32) return fail(c, MS_ERR_INVALID, "log_n > 32");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    ms_ntt_plan *p = (ms_ntt_plan *)calloc(1, sizeof *p);
    if (!p) return MS_ERR_NOMEM;
    p->ctx = c; p->field = field; p->direction = direction; p->log_n = log_n; p->offset = offset;
    *out = p;
    return MS_OK;
}
int ms_ntt_encode(ms_ntt_plan *p, void *column) {
    if (!p || !column) return MS_ERR_INVALID;
    if (p->ncols == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 16;
        p->cols = (void **)realloc(p->cols, p->cap * sizeof(void *));
    }
    p->cols[p->ncols++] = column;
    return MS_OK;
}
int ms_ntt_execute(ms_ntt_plan *p) {
    if (!p) return MS_ERR_INVALID;
    for (size_t k = 0; k < p->ncols; k++)
        orc_ntt_columns((u64 *)p->cols[k], 0, 1, (unsigned)p->field, p->log_n, p->offset, p->direction == MS_NTT_INVERSE);
    p->ncols = 0;
    return MS_OK;
}
int ms_ntt_plan_destroy(ms_ntt_plan *p) {
    if (p) { free(p->cols); free(p); }
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ pointwise stages */
int ms_pointwise(ms_ctx *c, int op, int df, void *dst, int lf, const void *lhs, int rf, const void *rhs, size_t n, size_t shift,
                 uint64_t exponent) {
    if (!c || !dst || !lhs) return MS_ERR_INVALID;
    if (bad_field(df) || bad_field(lf) || (rhs && bad_field(rf))) return fail(c, MS_ERR_INVALID, "unknown field id");
    if (op < MS_OP_MUL || op > MS_OP_SUB || op == MS_OP_FILL) return fail(c, MS_ERR_INVALID, "bad opcode %d", op);
    orc_pointwise(op, (unsigned)df, (u64 *)dst, (unsigned)lf, (const u64 *)lhs, (unsigned)rf, (const u64 *)rhs, n, shift, exponent);
    return MS_OK;
}
int ms_pointwise_const(ms_ctx *c, int op, int df, void *dst, int lf, const void *lhs, int cf, const uint64_t *cst, size_t n) {
    if (!c || !dst || !cst) return MS_ERR_INVALID;
    if (bad_field(df) || bad_field(cf) || (lhs && bad_field(lf))) return fail(c, MS_ERR_INVALID, "unknown field id");
    orc_pointwise_const(op, (unsigned)df, (u64 *)dst, (unsigned)lf, (const u64 *)lhs, (unsigned)cf, cst, n);
    return MS_OK;
}
int ms_sum_columns(ms_ctx *c, int field, const void *cols, size_t stride, unsigned ncols, size_t n, void *acc) {
    if (!c || !cols || !acc) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    orc_sum_columns((const u64 *)cols, stride * field, ncols, (unsigned)field, n, (u64 *)acc);
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ Merkle */
static int pow2_ge2(size_t n) { return n >= 2 && !(n & (n - 1)); }
int ms_hash_rows_sha256(ms_ctx *c, int field, const void *cols, size_t stride, unsigned ncols, size_t nrows, void *digests) {
    if (!c || !cols || !digests) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (ncols == 0) return fail(c, MS_ERR_INVALID, "ms_hash_rows_sha256: no columns");
    orc_hash_rows((const u64 *)cols, stride * field, ncols, (unsigned)field, nrows, (uint8_t *)digests);
    return MS_OK;
}
int ms_merkle_nodes_sha256(ms_ctx *c, const void *leaves, size_t n, void *nodes) {
    if (!c || !leaves || !nodes) return MS_ERR_INVALID;
    if (!pow2_ge2(n)) return fail(c, MS_ERR_INVALID, "merkle tree needs a power-of-two number of leaves >= 2, got %zu", n);
    orc_merkle_nodes((const uint8_t *)leaves, n, (uint8_t *)nodes);
    return MS_OK;
}
static int tree_over(ms_ctx *c, uint8_t *lv, size_t nrows, void *leaves, void *nodes, void *root) {
    uint8_t *nd = nodes ? (uint8_t *)nodes : (uint8_t *)malloc(nrows * 32);
    if (!nd) { if (!leaves) free(lv); return fail(c, MS_ERR_NOMEM, "out of host memory"); }
    orc_merkle_nodes(lv, nrows, nd);
    memcpy(root, nd + 32, 32);
    if (!nodes) free(nd);
    if (!leaves) free(lv);
    return MS_OK;
}
int ms_merkle_commit_sha256(ms_ctx *c, int field, const void *cols, size_t stride, unsigned ncols, size_t nrows, void *leaves,
                            void *nodes, void *root) {
    if (!c || !cols || !root) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (ncols == 0) return fail(c, MS_ERR_INVALID, "ms_merkle_commit_sha256: no columns");
    if (!pow2_ge2(nrows)) return fail(c, MS_ERR_INVALID, "merkle tree needs a power-of-two number of leaves >= 2, got %zu", nrows);
    uint8_t *lv = leaves ? (uint8_t *)leaves : (uint8_t *)malloc(nrows * 32);
    if (!lv) return fail(c, MS_ERR_NOMEM, "out of host memory");
    const double t0 = now_s();
    orc_hash_rows((const u64 *)cols, stride * field, ncols, (unsigned)field, nrows, lv);
    const int rc = tree_over(c, lv, nrows, leaves, nodes, root);
    return rc ? rc : done(c, "ms_merkle_commit", t0);
}
/* a FRI layer: rows of row_words consecutive words (src/fri.rs:199-216) */
int ms_merkle_commit_rows_sha256(ms_ctx *c, const void *rows, unsigned row_words, size_t nrows, void *leaves, void *nodes, void *root) {
    if (!c || !rows || !root || row_words == 0) return MS_ERR_INVALID;
    if (!pow2_ge2(nrows)) return fail(c, MS_ERR_INVALID, "merkle tree needs a power-of-two number of leaves >= 2, got %zu", nrows);
    uint8_t *lv = leaves ? (uint8_t *)leaves : (uint8_t *)malloc(nrows * 32);
    if (!lv) return fail(c, MS_ERR_NOMEM, "out of host memory");
    /* one "column" whose element is the whole row */
    const double t0 = now_s();
    orc_hash_rows((const u64 *)rows, 0, 1, row_words, nrows, lv);
    const int rc = tree_over(c, lv, nrows, leaves, nodes, root);
    return rc ? rc : done(c, "ms_merkle_commit_rows", t0);
}
int ms_pow_grind_sha256(ms_ctx *c, const uint8_t *seed, unsigned bits, uint64_t *nonce_out) {
    if (!c || !seed || !nonce_out) return MS_ERR_INVALID;
    if (bits > 64) return fail(c, MS_ERR_INVALID, "ms_pow_grind_sha256: at most 64 bits supported");
    const double t0 = now_s();
    *nonce_out = orc_pow_grind(seed, bits);
    return done(c, "ms_pow_grind", t0);
}

/* MerkleTreeImpl::prove (src/merkle.rs:149-207): sort + dedup, leaves paired with their siblings when both are asked
 * for, then the queue of parent indices walks up level by level */
static int cmp_u64(const void *a, const void *b) {
    const u64 x = *(const u64 *)a, y = *(const u64 *)b;
    return x < y ? -1 : x > y;
}
int ms_merkle_prove_sha256(ms_ctx *c, const void *leaves, const void *nodes, size_t n_leaves, const uint64_t *indices, unsigned n_indices,
                           uint8_t *initial_leaves, uint8_t *sibling_leaves, uint8_t *path_nodes, unsigned counts[3]) {
    if (!c || !leaves || !nodes || !indices || !initial_leaves || !sibling_leaves || !path_nodes || !counts) return MS_ERR_INVALID;
    if (!pow2_ge2(n_leaves)) return fail(c, MS_ERR_INVALID, "ms_merkle_prove: leaf count must be a power of two >= 2");
    for (unsigned k = 0; k < n_indices; k++)
        if (indices[k] >= n_leaves)
            return fail(c, MS_ERR_INVALID, "leaf index `%llu` cannot exceed the number of leaves (`%llu`)", (unsigned long long)indices[k],
                        (unsigned long long)n_leaves);
    unsigned height = 0;
    while (((size_t)1 << height) < n_leaves) height++;
    u64 *idx = (u64 *)malloc(sizeof(u64) * (n_indices + 1));
    u64 *queue = (u64 *)malloc(sizeof(u64) * ((size_t)n_indices * (height + 2) + 2));
    memcpy(idx, indices, sizeof(u64) * n_indices);
    qsort(idx, n_indices, sizeof(u64), cmp_u64);
    unsigned m = 0;
    for (unsigned k = 0; k < n_indices; k++)
        if (m == 0 || idx[m - 1] != idx[k]) idx[m++] = idx[k];
    const uint8_t *lv = (const uint8_t *)leaves, *nd = (const uint8_t *)nodes;
    size_t head = 0, tail = 0;
    unsigned ni = 0, ns = 0, np = 0;
    for (unsigned k = 0; k < m; k++) {
        const u64 i = idx[k];
        memcpy(initial_leaves + 32 * (size_t)ni++, lv + 32 * i, 32);
        queue[tail++] = (n_leaves + i) >> 1;
        if (k + 1 < m && idx[k + 1] == (i ^ 1)) {
            memcpy(initial_leaves + 32 * (size_t)ni++, lv + 32 * idx[++k], 32);
            continue;
        }
        memcpy(sibling_leaves + 32 * (size_t)ns++, lv + 32 * (i ^ 1), 32);
    }
    while (head < tail) {
        const u64 i = queue[head++];
        if (i > 2) queue[tail++] = i >> 1;
        if (head < tail && queue[head] == (i ^ 1)) { head++; continue; }
        memcpy(path_nodes + 32 * (size_t)np++, nd + 32 * (i ^ 1), 32);
    }
    counts[0] = ni; counts[1] = ns; counts[2] = np;
    free(idx);
    free(queue);
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ matrix plumbing */
int ms_matrix_from_rows(ms_ctx *c, int field, const void *rows, size_t n, unsigned k, void *cols, size_t stride) {
    if (!c || !rows || !cols) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (k > 1 && stride < n) return fail(c, MS_ERR_INVALID, "ms_matrix_from_rows: stride < n");
    const u64 *r = (const u64 *)rows;
    u64 *o = (u64 *)cols;
    const size_t f = (size_t)field;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++)
        for (unsigned j = 0; j < k; j++) memcpy(o + ((size_t)j * stride + i) * f, r + (i * k + j) * f, 8 * f);
    return MS_OK;
}
int ms_gather_rows(ms_ctx *c, int field, const void *cols, size_t stride, unsigned ncols, size_t nrows, const uint64_t *row_ids,
                   unsigned nq, void *out) {
    if (!c || !cols || !row_ids || !out) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    const size_t f = (size_t)field;
    for (unsigned q = 0; q < nq; q++) {
        if (row_ids[q] >= nrows) return fail(c, MS_ERR_INVALID, "ms_gather_rows: row %llu out of range", (unsigned long long)row_ids[q]);
        for (unsigned k = 0; k < ncols; k++)
            memcpy((u64 *)out + ((size_t)q * ncols + k) * f, (const u64 *)cols + ((size_t)k * stride + row_ids[q]) * f, 8 * f);
    }
    return MS_OK;
}
int ms_gather_rows_rowmajor(ms_ctx *c, const void *rows, unsigned row_words, size_t nrows, const uint64_t *row_ids, unsigned nq, void *out) {
    if (!c || !rows || !row_ids || !out) return MS_ERR_INVALID;
    for (unsigned q = 0; q < nq; q++) {
        if (row_ids[q] >= nrows) return fail(c, MS_ERR_INVALID, "ms_gather_rows_rowmajor: row %llu out of range", (unsigned long long)row_ids[q]);
        memcpy((u64 *)out + (size_t)q * row_words, (const u64 *)rows + row_ids[q] * row_words, 8 * (size_t)row_words);
    }
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ scan / FRI / DEEP */
int ms_scan_affine(ms_ctx *c, int field, const void *a, int a_field, const uint64_t *a_const, const void *b, int b_field, size_t n,
                   const uint64_t *init, int inclusive, void *out) {
    if (!c || !init || !out) return MS_ERR_INVALID;
    if (bad_field(field) || (a && bad_field(a_field)) || (b && bad_field(b_field))) return fail(c, MS_ERR_INVALID, "unknown field id");
    if ((a && a_field == MS_FIELD_FQ3 && field == MS_FIELD_FP) || (b && b_field == MS_FIELD_FQ3 && field == MS_FIELD_FP))
        return fail(c, MS_ERR_INVALID, "ms_scan_affine: Fq3 factors need an Fq3 result");
    if (!a && !a_const) return fail(c, MS_ERR_INVALID, "ms_scan_affine: neither a nor a_const");
    u64 init3[3] = {init[0], 0, 0}, ac3[3] = {0, 0, 0};
    if (field == MS_FIELD_FQ3) { init3[1] = init[1]; init3[2] = init[2]; }
    if (a_const) {
        ac3[0] = a_const[0];
        if (field == MS_FIELD_FQ3) { ac3[1] = a_const[1]; ac3[2] = a_const[2]; }
    }
    orc_scan_affine((unsigned)field, (const u64 *)a, (unsigned)a_field, a_const ? ac3 : NULL, (const u64 *)b, (unsigned)b_field, n, init3,
                    inclusive, (u64 *)out);
    return MS_OK;
}
int ms_fri_fold(ms_ctx *c, int field, const void *evals, unsigned log_n, unsigned log_ff, uint64_t offset, const uint64_t *alpha, void *out) {
    if (!c || !evals || !alpha || !out) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    if (log_ff == 0 || log_ff > 4 || log_ff > log_n) return fail(c, MS_ERR_INVALID, "ms_fri_fold: folding factor must be 2, 4, 8 or 16");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    const double t0 = now_s();
    orc_fri_apply_drp((const u64 *)evals, (unsigned)field, log_n, log_ff, offset, alpha, (u64 *)out);
    return done(c, "ms_fri_fold", t0);
}
/* horner_evaluate of every column at every point (src/composer.rs:43-86, src/utils.rs:124-131); the coefficient range is
 * split into blocks evaluated independently and recombined with powers of the point, so one column still uses all cores */
int ms_poly_eval(ms_ctx *c, int field, const void *coeffs, size_t stride, unsigned ncols, size_t n, const uint64_t *points,
                 unsigned npoints, uint64_t *out) {
    if (!c || !coeffs || !points || !out) return MS_ERR_INVALID;
    if (bad_field(field)) return fail(c, MS_ERR_INVALID, "unknown field id %d", field);
    const double t0 = now_s();
    const size_t BLK = 4096, nblk = (n + BLK - 1) / BLK;
    const size_t jobs = (size_t)ncols * npoints;
    fq3 *part = (fq3 *)malloc(sizeof(fq3) * (jobs * nblk + 1));
    if (!part) return fail(c, MS_ERR_NOMEM, "out of host memory");
    #pragma omp parallel for schedule(static)
    for (size_t t = 0; t < jobs * nblk; t++) {
        const size_t job = t / nblk, blk = t % nblk, col = job / npoints, k = job % npoints;
        const size_t s = blk * BLK, e = s + BLK < n ? s + BLK : n;
        orc_horner((const u64 *)coeffs + (col * stride + s) * field, (unsigned)field, e - s, points + 3 * k, part[t].c);
    }
    #pragma omp parallel for schedule(static)
    for (size_t job = 0; job < jobs; job++) {
        const fq3 x = load_el(points, 3, job % npoints), xb = fq3_pow(x, BLK);
        fq3 r = fq3_zero();
        for (size_t blk = nblk; blk-- > 0;) r = fq3_add(fq3_mul(r, xb), part[job * nblk + blk]);
        memcpy(out + 3 * job, &r, 24);
    }
    free(part);
    return done(c, "ms_poly_eval", t0);
}

/* one independent splitmix64 stream per word, rejecting draws >= p (the definition of ms_fill_random) */
int ms_fill_random(ms_ctx *c, void *dst, size_t nwords, uint64_t seed) {
    if (!c || !dst) return MS_ERR_INVALID;
    u64 *d = (u64 *)dst;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nwords; i++) {
        u64 s = seed ^ (0xD1B54A32D192ED03ULL * (u64)(i + 1)), z;
        do {
            s += 0x9E3779B97F4A7C15ULL;
            z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            z ^= z >> 31;
        } while (z >= GL_P);
        d[i] = fp_from_canon(z);
    }
    return MS_OK;
}

/* ------------------------------------------------------------------------------------------------ fused evaluator */
/* program words (ministark_b200/expr.py::compile_program): [op | a_is_fq << 8 | b_is_fq << 9, dst, a, b] */
enum { EV_X = 0, EV_CONST, EV_TRACE, EV_NEG, EV_ADD, EV_SUB, EV_MUL, EV_INV, EV_POW, EV_STORE, EV_PERIODIC };
#define EV_REGS 48
#define EV_CH 128   /* points per chunk: 48 registers x 3 lanes x 128 x 8 B = 144 KiB per thread, L2 resident */
typedef u64 evlane[EV_CH];

/* out[k] = 1 / in[k] over one chunk, zeros stay zero: one inversion + 3 multiplications per element */
static void chunk_inverse(int q, evlane *d, evlane *a, size_t cnt) {
    fq3 pref[EV_CH], acc = fq3_one();
    for (size_t k = 0; k < cnt; k++) {
        fq3 v = {{a[0][k], q ? a[1][k] : 0, q ? a[2][k] : 0}};
        pref[k] = acc;
        if (v.c[0] | v.c[1] | v.c[2]) acc = q ? fq3_mul(acc, v) : fq3_from_fp(fp_mul(acc.c[0], v.c[0]));
    }
    fq3 inv = q ? fq3_inv(acc) : fq3_from_fp(fp_inv(acc.c[0]));
    for (size_t k = cnt; k-- > 0;) {
        fq3 v = {{a[0][k], q ? a[1][k] : 0, q ? a[2][k] : 0}};
        if (!(v.c[0] | v.c[1] | v.c[2])) { d[0][k] = 0; if (q) d[1][k] = d[2][k] = 0; continue; }
        if (q) {
            fq3 r = fq3_mul(inv, pref[k]);
            inv = fq3_mul(inv, v);
            d[0][k] = r.c[0]; d[1][k] = r.c[1]; d[2][k] = r.c[2];
        } else {
            d[0][k] = fp_mul(inv.c[0], pref[k].c[0]);
            inv.c[0] = fp_mul(inv.c[0], v.c[0]);
        }
    }
}

static int eval_run(ms_ctx *c, const uint32_t *prog, unsigned nprog, const u64 *consts, unsigned nconsts, const u64 *const *cols,
                    const int *isq, unsigned ncols, int fq_field, unsigned log_m, u64 offset, int trace_bitrev, int out_bitrev, u64 *out) {
    /* validation, as the library does it: register file, constant pool, column table, defined-before-use */
    {
        char defined[EV_REGS] = {0};
        int stored = 0;
        for (unsigned k = 0; k < nprog; k++) {
            const uint32_t *ins = prog + 4 * k, op = ins[0] & 0xff;
            if (op > EV_PERIODIC || ins[1] >= EV_REGS) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: bad instruction %u", k);
            if (op == EV_CONST && ins[2] >= nconsts) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: constant index out of range");
            if (op == EV_TRACE || op == EV_PERIODIC) {
                if (ins[2] >= ncols) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: column %u out of range", ins[2]);
                if (isq[ins[2]] != (int)((ins[0] >> 8) & 1)) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: column %u has the wrong field", ins[2]);
                if (op == EV_PERIODIC && ins[3] > log_m) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: periodic table longer than the domain");
            }
            const int unary = op == EV_NEG || op == EV_INV || op == EV_POW || op == EV_STORE, binary = op == EV_ADD || op == EV_SUB || op == EV_MUL;
            if ((unary || binary) && (ins[2] >= EV_REGS || !defined[ins[2]]))
                return fail(c, MS_ERR_INVALID, "ms_eval_constraints: instruction %u reads register %u before it is written", k, ins[2]);
            if (binary && (ins[3] >= EV_REGS || !defined[ins[3]]))
                return fail(c, MS_ERR_INVALID, "ms_eval_constraints: instruction %u reads register %u before it is written", k, ins[3]);
            if (op == EV_STORE) stored = 1; else defined[ins[1]] = 1;
        }
        if (!stored) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: program stores no result");
    }
    const double t0 = now_s();
    const size_t M = (size_t)1 << log_m;
    const int fq3m = fq_field == 3;
    const unsigned fw = (unsigned)fq_field;
    /* x = offset * g^i from a two-level table */
    const size_t lo_len = M < 4096 ? M : 4096, hi_len = (M + 4095) / 4096;
    u64 *lo = (u64 *)malloc(8 * lo_len), *hi = (u64 *)malloc(8 * hi_len);
    const u64 g = orc_root_of_unity(log_m);
    lo[0] = GL_ONE;
    for (size_t e = 1; e < lo_len; e++) lo[e] = fp_mul(lo[e - 1], g);
    const u64 g_hi = fp_pow(g, 4096);
    hi[0] = GL_ONE;
    for (size_t e = 1; e < hi_len; e++) hi[e] = fp_mul(hi[e - 1], g_hi);
    out_bitrev = out_bitrev && trace_bitrev;
    const size_t nchunks = (M + EV_CH - 1) / EV_CH;
    #pragma omp parallel
    {
        evlane(*r)[3] = (evlane(*)[3])malloc(sizeof(evlane) * 3 * EV_REGS);
        size_t pt[EV_CH];
        #pragma omp for schedule(static)
        for (size_t ch = 0; ch < nchunks; ch++) {
            const size_t t0 = ch * EV_CH, cnt = t0 + EV_CH <= M ? EV_CH : M - t0;
            for (size_t k = 0; k < cnt; k++) pt[k] = (trace_bitrev && log_m) ? bitrev(t0 + k, log_m) : t0 + k;
            for (unsigned pc = 0; pc < nprog; pc++) {
                const uint32_t *ins = prog + 4 * pc, op = ins[0] & 0xff, d = ins[1], a = ins[2], b = ins[3];
                const int colq = (ins[0] >> 8) & 1, qa = colq && fq3m, qb = ((ins[0] >> 9) & 1) && fq3m;
                switch (op) {
                case EV_X:
                    for (size_t k = 0; k < cnt; k++) r[d][0][k] = fp_mul(fp_mul(hi[pt[k] >> 12], lo[pt[k] & 4095]), offset);
                    break;
                case EV_CONST:
                    for (size_t k = 0; k < cnt; k++) {
                        r[d][0][k] = consts[3 * (size_t)a];
                        if (qa) { r[d][1][k] = consts[3 * (size_t)a + 1]; r[d][2][k] = consts[3 * (size_t)a + 2]; }
                    }
                    break;
                case EV_TRACE:
                case EV_PERIODIC:
                    for (size_t k = 0; k < cnt; k++) {
                        size_t pos;
                        if (op == EV_TRACE) {
                            pos = (pt[k] + (size_t)b) & (M - 1);
                            if (trace_bitrev && log_m) pos = bitrev(pos, log_m);
                        } else {
                            pos = pt[k] & (((size_t)1 << b) - 1);
                        }
                        if (colq) {
                            const u64 *e = cols[a] + pos * fw;
                            r[d][0][k] = e[0];
                            if (fq3m) { r[d][1][k] = e[1]; r[d][2][k] = e[2]; }
                        } else {
                            r[d][0][k] = cols[a][pos];
                        }
                    }
                    break;
                case EV_NEG:
                    for (int l = 0; l < (qa ? 3 : 1); l++)
                        for (size_t k = 0; k < cnt; k++) r[d][l][k] = fp_neg(r[a][l][k]);
                    break;
                case EV_ADD:
                case EV_SUB:
                    for (int l = 0; l < ((qa || qb) ? 3 : 1); l++)
                        for (size_t k = 0; k < cnt; k++) {
                            const u64 x = (l == 0 || qa) ? r[a][l][k] : 0, y = (l == 0 || qb) ? r[b][l][k] : 0;
                            r[d][l][k] = op == EV_ADD ? fp_add(x, y) : fp_sub(x, y);
                        }
                    break;
                case EV_MUL:
                    if (!qa && !qb) {
                        for (size_t k = 0; k < cnt; k++) r[d][0][k] = fp_mul(r[a][0][k], r[b][0][k]);
                    } else if (qa && qb) {
                        for (size_t k = 0; k < cnt; k++) {
                            const fq3 x = {{r[a][0][k], r[a][1][k], r[a][2][k]}}, y = {{r[b][0][k], r[b][1][k], r[b][2][k]}}, z = fq3_mul(x, y);
                            r[d][0][k] = z.c[0]; r[d][1][k] = z.c[1]; r[d][2][k] = z.c[2];
                        }
                    } else {
                        const uint32_t q = qa ? a : b, s = qa ? b : a;
                        for (size_t k = 0; k < cnt; k++) {
                            const u64 y = r[s][0][k], x0 = r[q][0][k], x1 = r[q][1][k], x2 = r[q][2][k];
                            r[d][0][k] = fp_mul(x0, y); r[d][1][k] = fp_mul(x1, y); r[d][2][k] = fp_mul(x2, y);
                        }
                    }
                    break;
                case EV_INV:
                    if (d == a) {
                        evlane tmp[3];
                        memcpy(tmp, r[a], sizeof tmp);
                        chunk_inverse(qa, r[d], tmp, cnt);
                    } else {
                        chunk_inverse(qa, r[d], r[a], cnt);
                    }
                    break;
                case EV_POW:
                    for (size_t k = 0; k < cnt; k++) {
                        if (qa) {
                            const fq3 x = {{r[a][0][k], r[a][1][k], r[a][2][k]}}, z = fq3_pow(x, b);
                            r[d][0][k] = z.c[0]; r[d][1][k] = z.c[1]; r[d][2][k] = z.c[2];
                        } else {
                            r[d][0][k] = fp_pow(r[a][0][k], b);
                        }
                    }
                    break;
                case EV_STORE:
                    for (size_t k = 0; k < cnt; k++) {
                        u64 *o = out + (out_bitrev ? t0 + k : pt[k]) * fw;
                        o[0] = r[a][0][k];
                        if (fq3m) { o[1] = qa ? r[a][1][k] : 0; o[2] = qa ? r[a][2][k] : 0; }
                    }
                    break;
                default: break;
                }
            }
        }
        free(r);
    }
    free(lo);
    free(hi);
    return done(c, "ms_eval_constraints", t0);
}

int ms_eval_constraints(ms_ctx *c, const uint32_t *program, unsigned nprog, const uint64_t *consts, unsigned nconsts, const void *base_cols,
                        size_t base_stride, unsigned nbase, const void *ext_cols, size_t ext_stride, unsigned next, int fq_field,
                        unsigned log_m, uint64_t offset, int trace_bitrev, int out_bitrev, void *out) {
    if (!c || !program || !consts || !out || nprog == 0) return MS_ERR_INVALID;
    if (bad_field(fq_field)) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: bad Fq field id");
    if (log_m > 32) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: domain too large");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    if ((nbase && !base_cols) || (next && !ext_cols)) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: missing columns");
    const size_t M = (size_t)1 << log_m;
    if ((nbase > 1 && base_stride < M) || (next > 1 && ext_stride < M)) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: stride < domain");
    const unsigned ncols = nbase + next;
    const u64 **cols = (const u64 **)malloc(sizeof(u64 *) * (ncols ? ncols : 1));
    int *isq = (int *)malloc(sizeof(int) * (ncols ? ncols : 1));
    for (unsigned i = 0; i < nbase; i++) { cols[i] = (const u64 *)base_cols + (size_t)i * base_stride; isq[i] = 0; }
    for (unsigned i = 0; i < next; i++) { cols[nbase + i] = (const u64 *)ext_cols + (size_t)i * ext_stride * fq_field; isq[nbase + i] = 1; }
    rc = eval_run(c, program, nprog, consts, nconsts, cols, isq, ncols, fq_field, log_m, offset, trace_bitrev, out_bitrev, (u64 *)out);
    free(cols);
    free(isq);
    return rc;
}
int ms_eval_constraints_ptrs(ms_ctx *c, const uint32_t *program, unsigned nprog, const uint64_t *consts, unsigned nconsts,
                             const void *const *col_ptrs, const int *col_is_fq, unsigned ncols, int fq_field, unsigned log_m,
                             uint64_t offset, int trace_bitrev, int out_bitrev, void *out) {
    if (!c || !program || !consts || !out || nprog == 0 || (ncols && (!col_ptrs || !col_is_fq))) return MS_ERR_INVALID;
    if (bad_field(fq_field)) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: bad Fq field id");
    if (log_m > 32) return fail(c, MS_ERR_INVALID, "ms_eval_constraints: domain too large");
    int rc = check_offset(c, offset);
    if (rc) return rc;
    int *isq = (int *)malloc(sizeof(int) * (ncols ? ncols : 1));
    for (unsigned i = 0; i < ncols; i++) {
        if (!col_ptrs[i]) { free(isq); return fail(c, MS_ERR_INVALID, "ms_eval_constraints_ptrs: column %u is NULL", i); }
        isq[i] = col_is_fq[i] ? 1 : 0;
    }
    rc = eval_run(c, program, nprog, consts, nconsts, (const u64 *const *)col_ptrs, isq, ncols, fq_field, log_m, offset, trace_bitrev,
                  out_bitrev, (u64 *)out);
    free(isq);
    return rc;
}
