// pointwise.cu — element-wise field stages.
//
// One CUDA kernel family covers the reference's 17 Metal templates / 58 instantiations
// (gpu/src/metal/evaluation_shaders.h.metal:11-168, wrapped by 14 *Stage structs in
// gpu/src/stage.rs): MulInto/MulAssign (+Const), AddInto/AddAssign (+Const), ConvertInto,
// InverseInto/InPlace, ExpInto/InPlace, NegInto/InPlace, MulPow, FillBuff; operand fields
// Fp x Fp, Fq3 x Fp, Fq3 x Fq3.  rhs is read at (i + shift) % n as in the reference.
// Also Matrix::sum_columns (src/matrix.rs:322-394) as ONE pass over all columns instead of
// one AddAssign dispatch per column.
// These are streaming kernels (HBM bound): one element per thread, grid sized to cover n.
#include "ctx.cuh"
#include <vector>

namespace ms {

using gl::Fq3;

template <int F>
struct Elem;
template <>
struct Elem<1> {
    typedef u64 T;
    static __device__ __forceinline__ T load(const u64 *p, size_t i) { return p[i]; }
    static __device__ __forceinline__ void store(u64 *p, size_t i, T v) { p[i] = v; }
};
template <>
struct Elem<3> {
    typedef Fq3 T;
    static __device__ __forceinline__ T load(const u64 *p, size_t i) { return Fq3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
    static __device__ __forceinline__ void store(u64 *p, size_t i, T v) {
        p[3 * i] = v.c0; p[3 * i + 1] = v.c1; p[3 * i + 2] = v.c2;
    }
};

__device__ __forceinline__ Fq3 lift(u64 a) { return gl::fq3(a); }
__device__ __forceinline__ Fq3 lift(Fq3 a) { return a; }
// mixed products pick the cheapest form
__device__ __forceinline__ u64 fmul(u64 a, u64 b) { return gl::mul(a, b); }
__device__ __forceinline__ Fq3 fmul(Fq3 a, u64 b) { return gl::mul(a, b); }
__device__ __forceinline__ Fq3 fmul(u64 a, Fq3 b) { return gl::mul(b, a); }
__device__ __forceinline__ Fq3 fmul(Fq3 a, Fq3 b) { return gl::mul(a, b); }
__device__ __forceinline__ u64 fadd(u64 a, u64 b) { return gl::add(a, b); }
__device__ __forceinline__ Fq3 fadd(Fq3 a, u64 b) { return Fq3{gl::add(a.c0, b), a.c1, a.c2}; }
__device__ __forceinline__ Fq3 fadd(u64 a, Fq3 b) { return Fq3{gl::add(a, b.c0), b.c1, b.c2}; }
__device__ __forceinline__ Fq3 fadd(Fq3 a, Fq3 b) { return gl::add(a, b); }
__device__ __forceinline__ u64 fsub(u64 a, u64 b) { return gl::sub(a, b); }
__device__ __forceinline__ Fq3 fsub(Fq3 a, u64 b) { return Fq3{gl::sub(a.c0, b), a.c1, a.c2}; }
__device__ __forceinline__ Fq3 fsub(u64 a, Fq3 b) { return Fq3{gl::sub(a, b.c0), gl::neg(b.c1), gl::neg(b.c2)}; }
__device__ __forceinline__ Fq3 fsub(Fq3 a, Fq3 b) { return gl::sub(a, b); }
__device__ __forceinline__ u64 finv(u64 a) { return gl::inv(a); }
__device__ __forceinline__ Fq3 finv(Fq3 a) { return gl::inv(a); }
__device__ __forceinline__ u64 fpow(u64 a, u64 e) { return gl::pow(a, e); }
__device__ __forceinline__ Fq3 fpow(Fq3 a, u64 e) { return gl::pow(a, e); }
__device__ __forceinline__ u64 fneg(u64 a) { return gl::neg(a); }
__device__ __forceinline__ Fq3 fneg(Fq3 a) { return gl::neg(a); }

// store a value of type V into a destination of field DF (DF >= field of V, or V has c1=c2=0)
template <int DF>
__device__ __forceinline__ void put(u64 *dst, size_t i, u64 v) {
    if (DF == 1) dst[i] = v;
    else Elem<3>::store(dst, i, gl::fq3(v));
}
template <int DF>
__device__ __forceinline__ void put(u64 *dst, size_t i, Fq3 v) {
    if (DF == 1) dst[i] = v.c0;
    else Elem<3>::store(dst, i, v);
}

template <int DF, int LF, int RF>
__global__ void pointwise_kernel(int op, u64 *dst, const u64 *lhs, const u64 *rhs, size_t n, size_t shift, u64 exponent) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Elem<LF>::T a = Elem<LF>::load(lhs, i);
    size_t j = i + shift;
    if (j >= n) j -= n;
    switch (op) {
        case MS_OP_MUL: put<DF>(dst, i, fmul(a, Elem<RF>::load(rhs, j))); break;
        case MS_OP_ADD: put<DF>(dst, i, fadd(a, Elem<RF>::load(rhs, j))); break;
        case MS_OP_SUB: put<DF>(dst, i, fsub(a, Elem<RF>::load(rhs, j))); break;
        case MS_OP_MULPOW: put<DF>(dst, i, fmul(a, fpow(Elem<RF>::load(rhs, j), exponent))); break;
        case MS_OP_CONVERT: put<DF>(dst, i, a); break;
        case MS_OP_INV: put<DF>(dst, i, finv(a)); break;
        case MS_OP_EXP: put<DF>(dst, i, fpow(a, exponent)); break;
        case MS_OP_NEG: put<DF>(dst, i, fneg(a)); break;
        default: break;
    }
}

template <int DF, int LF, int CF>
__global__ void pointwise_const_kernel(int op, u64 *dst, const u64 *lhs, u64 k0, u64 k1, u64 k2, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Elem<CF>::T k;
    if constexpr (CF == 1) k = k0;
    else k = Fq3{k0, k1, k2};
    if (op == MS_OP_FILL) {
        put<DF>(dst, i, k);
        return;
    }
    typename Elem<LF>::T a = Elem<LF>::load(lhs, i);
    switch (op) {
        case MS_OP_MUL: put<DF>(dst, i, fmul(a, k)); break;
        case MS_OP_ADD: put<DF>(dst, i, fadd(a, k)); break;
        case MS_OP_SUB: put<DF>(dst, i, fsub(a, k)); break;
        default: break;
    }
}

__global__ void sum_columns_kernel(const u64 *cols, size_t col_stride_words, unsigned ncols, size_t nwords, u64 *acc) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    u64 s = 0;
    for (unsigned c = 0; c < ncols; c++) s = gl::add(s, cols[(size_t)c * col_stride_words + i]);
    acc[i] = s;
}

// dst column c, row i  <-  src row i, entry c   (row-major [n][k] -> column-major [k][n], elements of EW words)
__global__ void from_rows_kernel(const u64 *src, u64 *dst, size_t n, unsigned k, unsigned ew, size_t dst_col_stride_words) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;   // over n * k elements, column-major order
    if (idx >= n * k) return;
    const size_t i = idx % n, c = idx / n;
    for (unsigned w = 0; w < ew; w++) dst[c * dst_col_stride_words + i * ew + w] = src[(i * k + c) * ew + w];
}
// out[q][c] = cols[c][rows[q]]
__global__ void gather_rows_kernel(const u64 *cols, size_t col_stride_words, unsigned ncols, unsigned ew, const u64 *rows,
                                   unsigned nq, u64 *out) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * ncols) return;
    const unsigned q = idx / ncols, cc = idx % ncols;
    for (unsigned w = 0; w < ew; w++) out[((size_t)q * ncols + cc) * ew + w] = cols[(size_t)cc * col_stride_words + rows[q] * ew + w];
}

// diagnostic: the lazy field primitives the NTT butterflies are built from, applied element-wise (tests only)
__global__ void lazy_ops_kernel(const u64 *a, const u64 *b, size_t n, u64 *out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 x = a[i], y = b[i], yc = gl::canon(y);
    out[i] = gl::add_lc(x, yc);
    out[n + i] = gl::sub_lc(x, yc);
    out[2 * n + i] = gl::add_ll(x, y);
    out[3 * n + i] = gl::sub_ll(x, y);
    out[4 * n + i] = gl::mul(x, yc);
}

}  // namespace ms

using namespace ms;

extern "C" {

int ms_pointwise(ms_ctx *c, int op, int df, void *dst, int lf, const void *lhs, int rf, const void *rhs, size_t n,
                 size_t shift, uint64_t exponent) {
    if (!c || !dst || !lhs) return MS_ERR_INVALID;
    const bool binary = (op == MS_OP_MUL || op == MS_OP_ADD || op == MS_OP_SUB || op == MS_OP_MULPOW);
    const bool unary = (op == MS_OP_CONVERT || op == MS_OP_INV || op == MS_OP_EXP || op == MS_OP_NEG);
    if (!binary && !unary) return fail(c, MS_ERR_INVALID, "ms_pointwise: unknown op %d", op);
    if ((df != 1 && df != 3) || (lf != 1 && lf != 3)) return fail(c, MS_ERR_INVALID, "ms_pointwise: bad field id");
    if (binary && (!rhs || (rf != 1 && rf != 3))) return fail(c, MS_ERR_INVALID, "ms_pointwise: binary op needs rhs");
    if (!binary) rf = 1;
    const int res_field = (lf == 3 || (binary && rf == 3)) ? 3 : 1;
    if (df < res_field) return fail(c, MS_ERR_INVALID, "ms_pointwise: destination field too small for the result");
    if (n == 0) return MS_OK;
    if (shift >= n) shift %= n;
    const bool alias = (dst == lhs);
    Staged L(c, lhs, n * lf * 8, true, alias);
    if (L.rc) return L.rc;
    Staged Rr(c, binary ? rhs : nullptr, binary ? n * rf * 8 : 0, true, false);
    if (Rr.rc) return Rr.rc;
    Staged D(c, alias ? nullptr : dst, alias ? 0 : n * df * 8, false, true);
    if (D.rc) return D.rc;
    u64 *d = alias ? L.as<u64>() : D.as<u64>();
    const u64 *l = L.as<u64>(), *r = binary ? Rr.as<u64>() : l;
    if (alias && df != lf) return fail(c, MS_ERR_INVALID, "ms_pointwise: in-place form needs dst field == lhs field");
    // in-place with a shifted rhs that aliases dst would race (the reference never does that)
    const unsigned threads = 256, blocks = (unsigned)((n + threads - 1) / threads);
#define MS_PW(DF, LF, RF) pointwise_kernel<DF, LF, RF><<<blocks, threads, 0, c->stream>>>(op, d, l, r, n, shift, exponent)
    if (df == 1) MS_PW(1, 1, 1);
    else if (lf == 1 && rf == 1) MS_PW(3, 1, 1);
    else if (lf == 1 && rf == 3) MS_PW(3, 1, 3);
    else if (lf == 3 && rf == 1) MS_PW(3, 3, 1);
    else MS_PW(3, 3, 3);
#undef MS_PW
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = L.finish())) return rc;
    if ((rc = Rr.finish())) return rc;
    return D.finish();
}

int ms_pointwise_const(ms_ctx *c, int op, int df, void *dst, int lf, const void *lhs, int cf, const uint64_t *k, size_t n) {
    if (!c || !dst || !k) return MS_ERR_INVALID;
    if (op != MS_OP_MUL && op != MS_OP_ADD && op != MS_OP_SUB && op != MS_OP_FILL)
        return fail(c, MS_ERR_INVALID, "ms_pointwise_const: unknown op %d", op);
    if ((df != 1 && df != 3) || (cf != 1 && cf != 3)) return fail(c, MS_ERR_INVALID, "ms_pointwise_const: bad field id");
    if (op == MS_OP_FILL) { lf = 1; lhs = nullptr; }
    else if (!lhs || (lf != 1 && lf != 3)) return fail(c, MS_ERR_INVALID, "ms_pointwise_const: needs lhs");
    const int res_field = (lf == 3 || cf == 3) ? 3 : 1;
    if (df < res_field) return fail(c, MS_ERR_INVALID, "ms_pointwise_const: destination field too small");
    if (n == 0) return MS_OK;
    u64 kk[3] = {0, 0, 0};
    if (is_device_ptr(k)) MS_CUDA(c, cudaMemcpy(kk, k, cf * 8, cudaMemcpyDeviceToHost));
    else for (int i = 0; i < cf; i++) kk[i] = k[i];
    const bool alias = (dst == lhs);
    if (alias && df != lf) return fail(c, MS_ERR_INVALID, "ms_pointwise_const: in-place form needs dst field == lhs field");
    Staged L(c, lhs, lhs ? n * lf * 8 : 0, true, alias);
    if (L.rc) return L.rc;
    Staged D(c, alias ? nullptr : dst, alias ? 0 : n * df * 8, false, true);
    if (D.rc) return D.rc;
    u64 *d = alias ? L.as<u64>() : D.as<u64>();
    const u64 *l = L.as<u64>();
    const unsigned threads = 256, blocks = (unsigned)((n + threads - 1) / threads);
#define MS_PC(DF, LF, CF) pointwise_const_kernel<DF, LF, CF><<<blocks, threads, 0, c->stream>>>(op, d, l, kk[0], kk[1], kk[2], n)
    if (df == 1) MS_PC(1, 1, 1);
    else if (lf == 1 && cf == 1) MS_PC(3, 1, 1);
    else if (lf == 1 && cf == 3) MS_PC(3, 1, 3);
    else if (lf == 3 && cf == 1) MS_PC(3, 3, 1);
    else MS_PC(3, 3, 3);
#undef MS_PC
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = L.finish())) return rc;
    return D.finish();
}

int ms_sum_columns(ms_ctx *c, int field, const void *cols, size_t col_stride_elems, unsigned ncols, size_t n, void *acc) {
    if (!c || !cols || !acc) return MS_ERR_INVALID;
    if (field != 1 && field != 3) return fail(c, MS_ERR_INVALID, "ms_sum_columns: bad field id");
    if (ncols > 1 && col_stride_elems < n) return fail(c, MS_ERR_INVALID, "ms_sum_columns: stride < n");
    if (n == 0) return MS_OK;
    Staged in(c, cols, ncols ? ((size_t)(ncols - 1) * col_stride_elems + n) * field * 8 : 0, true, false);
    if (in.rc) return in.rc;
    Staged out(c, acc, n * field * 8, false, true);
    if (out.rc) return out.rc;
    const size_t nwords = n * field;
    sum_columns_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, c->stream>>>(in.as<u64>(), col_stride_elems * field, ncols,
                                                                                 nwords, out.as<u64>());
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = in.finish())) return rc;
    return out.finish();
}


// Matrix::from_arrays / from_rows (src/matrix.rs:33-64) and the composition-polynomial split of
// src/prover.rs:113-120: n rows of k elements (row-major) -> k columns of n elements.
int ms_matrix_from_rows(ms_ctx *c, int field, const void *rows, size_t n, unsigned k, void *cols, size_t col_stride_elems) {
    if (!c || !rows || !cols || k == 0) return MS_ERR_INVALID;
    if (field != 1 && field != 3) return fail(c, MS_ERR_INVALID, "ms_matrix_from_rows: bad field id");
    if (k > 1 && col_stride_elems < n) return fail(c, MS_ERR_INVALID, "ms_matrix_from_rows: stride < n");
    if (n == 0) return MS_OK;
    Staged in(c, rows, n * k * field * 8, true, false);
    if (in.rc) return in.rc;
    Staged out(c, cols, ((size_t)(k - 1) * col_stride_elems + n) * field * 8, false, true);
    if (out.rc) return out.rc;
    const size_t total = n * k;
    from_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(in.as<u64>(), out.as<u64>(), n, k, (unsigned)field,
                                                                              col_stride_elems * field);
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = in.finish())) return rc;
    return out.finish();
}

// Matrix::get_row for a list of rows (src/matrix.rs:288-294; Queries::new src/trace.rs:115-157, query_layer
// src/fri.rs:650-664): out[q * ncols + c] = cols[c][row_ids[q]].  row_ids: host array.
int ms_gather_rows(ms_ctx *c, int field, const void *cols, size_t col_stride_elems, unsigned ncols, size_t nrows,
                   const uint64_t *row_ids, unsigned nq, void *out) {
    if (!c || !cols || !row_ids || !out) return MS_ERR_INVALID;
    if (field != 1 && field != 3) return fail(c, MS_ERR_INVALID, "ms_gather_rows: bad field id");
    if (ncols == 0 || nq == 0) return MS_OK;
    for (unsigned q = 0; q < nq; q++)
        if (row_ids[q] >= nrows) return fail(c, MS_ERR_INVALID, "ms_gather_rows: row %llu out of range", (unsigned long long)row_ids[q]);
    if (ncols > 1 && col_stride_elems < nrows) return fail(c, MS_ERR_INVALID, "ms_gather_rows: stride < nrows");
    Staged in(c, cols, ((size_t)(ncols - 1) * col_stride_elems + nrows) * field * 8, true, false);
    if (in.rc) return in.rc;
    Staged o(c, out, (size_t)nq * ncols * field * 8, false, true);
    if (o.rc) return o.rc;
    void *ids;
    int rc = scratch_get(c, 3, (size_t)nq * 8, &ids);
    if (rc) return rc;
    MS_CUDA(c, cudaMemcpyAsync(ids, row_ids, (size_t)nq * 8, cudaMemcpyHostToDevice, c->stream));
    const unsigned total = nq * ncols;
    gather_rows_kernel<<<(total + 127) / 128, 128, 0, c->stream>>>(in.as<u64>(), col_stride_elems * field, ncols, (unsigned)field,
                                                                    (const u64 *)ids, nq, o.as<u64>());
    c->launches++;
    MS_CHECK_LAUNCH(c);
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    if ((rc = in.finish())) return rc;
    return o.finish();
}

// query_layer (src/fri.rs:650-664): rows of a committed FRI layer, which is kept ROW-MAJOR (row k = the ff consecutive
// evaluations of coset k, as ms_merkle_commit_rows_sha256 hashed them).  out[q] = rows[row_ids[q]] (row_words words each).
int ms_gather_rows_rowmajor(ms_ctx *c, const void *rows, unsigned row_words, size_t nrows, const uint64_t *row_ids, unsigned nq,
                            void *out) {
    if (!c || !rows || !row_ids || !out) return MS_ERR_INVALID;
    if (row_words == 0 || nq == 0) return MS_OK;
    std::vector<u64> first(nq);
    for (unsigned q = 0; q < nq; q++) {
        if (row_ids[q] >= nrows) return fail(c, MS_ERR_INVALID, "ms_gather_rows_rowmajor: row %llu out of range", (unsigned long long)row_ids[q]);
        first[q] = row_ids[q] * row_words;
    }
    Staged in(c, rows, nrows * row_words * 8, true, false);
    if (in.rc) return in.rc;
    Staged o(c, out, (size_t)nq * row_words * 8, false, true);
    if (o.rc) return o.rc;
    void *ids;
    int rc = scratch_get(c, 3, (size_t)nq * 8, &ids);
    if (rc) return rc;
    MS_CUDA(c, cudaMemcpyAsync(ids, first.data(), (size_t)nq * 8, cudaMemcpyHostToDevice, c->stream));
    // word (q, w) = rows[first[q] + w]: the column-major gather with one-word "columns" of stride 1
    const unsigned total = nq * row_words;
    gather_rows_kernel<<<(total + 127) / 128, 128, 0, c->stream>>>(in.as<u64>(), 1, row_words, 1u, (const u64 *)ids, nq, o.as<u64>());
    c->launches++;
    MS_CHECK_LAUNCH(c);
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    if ((rc = in.finish())) return rc;
    return o.finish();
}

// Diagnostic (tests only): out[k*n + i] = op_k(a[i], b[i]) for the lazy primitives of field.cuh — k = 0 add_lc(a, canon b),
// 1 sub_lc(a, canon b), 2 add_ll(a, b), 3 sub_ll(a, b), 4 mul(a, canon b).  a, b: ANY 64-bit words.  Results of 0..3 are
// lazy (any u64 congruent to the exact result mod p); mul is canonical.
int ms_debug_lazy_ops(ms_ctx *c, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
    if (!c || !a || !b || !out) return MS_ERR_INVALID;
    if (n == 0) return MS_OK;
    Staged A(c, a, n * 8, true, false);
    if (A.rc) return A.rc;
    Staged B(c, b, n * 8, true, false);
    if (B.rc) return B.rc;
    Staged O(c, out, 5 * n * 8, false, true);
    if (O.rc) return O.rc;
    lazy_ops_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(A.as<u64>(), B.as<u64>(), n, O.as<u64>());
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = A.finish())) return rc;
    if ((rc = B.finish())) return rc;
    return O.finish();
}

}  // extern "C"
