// deep.cu — out-of-domain evaluation of coefficient-form polynomials.
//
// DeepPolyComposer::get_ood_evals (src/composer.rs:43-86) evaluates every trace / composition
// polynomial at z * g^offset with Horner's rule (horner_evaluate, src/utils.rs:124-131): a sequential
// recurrence over n coefficients per (column, point), parallel only across columns on the CPU.
// Here one evaluation is a two-level parallel reduction:
//     P(x) = sum_b x^(b*S) * sum_t x^t * sum_m c[b*S + t + 256 m] * (x^256)^m
// thread t of block b runs Horner in y = x^256 over a coalesced stride-256 slice, the block sums
// x^t-weighted terms in shared memory, and the per-block partials form a new (short) polynomial in
// x^S that is reduced by the same kernel until one value is left.  Each column is read once per call
// for ALL requested points.  The DEEP quotients themselves are evaluated pointwise on the LDE
// (ministark_b200/deep.py builds the expression; csrc/eval.cu runs it), which equals the reference's
// coefficient-form synthetic division (src/composer.rs:89-188) followed by the LDE, value for value.
#include "ctx.cuh"

namespace ms {

using gl::Fq3;

constexpr int kPeThreads = 256;
constexpr int kPeChunk = 64;                      // coefficients per thread
constexpr size_t kPeSpan = (size_t)kPeThreads * kPeChunk;

template <int F>
__device__ __forceinline__ Fq3 load_coeff(const u64 *p, size_t i) {
    if (F == 1) return gl::fq3(p[i]);
    return Fq3{p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}

// partials[col * out_col_stride + blockIdx.z * nblocks + b] = sum over the block's coefficient range, relative to
// its start, evaluated at points[blockIdx.z]
template <int F>
__global__ void __launch_bounds__(kPeThreads) poly_eval_kernel(const u64 *__restrict__ coeffs, size_t col_stride_words,
                                                                size_t n, const u64 *__restrict__ points,
                                                                u64 *__restrict__ partials, size_t out_col_stride) {
    __shared__ Fq3 red[kPeThreads];
    const unsigned t = threadIdx.x, b = blockIdx.x, col = blockIdx.y, k = blockIdx.z;
    const u64 *c = coeffs + (size_t)col * col_stride_words;
    const Fq3 x{points[3 * k], points[3 * k + 1], points[3 * k + 2]};
    Fq3 y = x;                                    // y = x^256
#pragma unroll 1
    for (int i = 0; i < 8; i++) y = gl::sqr(y);
    const size_t start = (size_t)b * kPeSpan, end = min(n, start + kPeSpan);
    Fq3 acc = gl::fq3(0);
#pragma unroll 1
    for (int m = kPeChunk - 1; m >= 0; m--) {
        const size_t idx = start + t + (size_t)kPeThreads * m;
        acc = gl::mul(acc, y);
        if (idx < end) acc = gl::add(acc, load_coeff<F>(c, idx));
    }
    red[t] = gl::mul(acc, gl::pow(x, (u64)t));
    __syncthreads();
    for (int s = kPeThreads / 2; s > 0; s >>= 1) {
        if ((int)t < s) red[t] = gl::add(red[t], red[t + s]);
        __syncthreads();
    }
    if (t == 0) {
        u64 *o = partials + 3 * ((size_t)col * out_col_stride + (size_t)k * gridDim.x + b);
        o[0] = red[0].c0; o[1] = red[0].c1; o[2] = red[0].c2;
    }
}

}  // namespace ms

using namespace ms;

// out[(col * npoints + k) * 3 .. +3] = P_col(points[k])   (Fq3 values; points are Fq3 elements)
extern "C" int ms_poly_eval(ms_ctx *c, int field, const void *coeffs, size_t col_stride_elems, unsigned ncols, size_t n,
                            const uint64_t *points, unsigned npoints, uint64_t *out) {
    if (!c || !coeffs || !points || !out) return MS_ERR_INVALID;
    if (field != 1 && field != 3) return fail(c, MS_ERR_INVALID, "ms_poly_eval: bad field id");
    if (ncols == 0 || npoints == 0 || n == 0 || npoints > 65535 || ncols > 65535) return fail(c, MS_ERR_INVALID, "ms_poly_eval: bad sizes");
    if (ncols > 1 && col_stride_elems < n) return fail(c, MS_ERR_INVALID, "ms_poly_eval: stride < n");
    Staged in(c, coeffs, ((size_t)(ncols - 1) * col_stride_elems + n) * field * 8, true, false);
    if (in.rc) return in.rc;
    const unsigned K = npoints;
    std::vector<u64> pts(3 * (size_t)K);   // host copy of the points (tiny): each level uses point^span
    MS_CUDA(c, cudaMemcpy(pts.data(), points, pts.size() * 8, cudaMemcpyDefault));
    const size_t nb0 = (n + kPeSpan - 1) / kPeSpan;
    const size_t lvl_words = 3 * (size_t)ncols * K * nb0;
    void *scr;
    int rc = scratch_get(c, 3, (2 * lvl_words + 3 * (size_t)K) * 8 + 256, &scr);
    if (rc) return rc;
    u64 *cur = (u64 *)scr, *nxt = cur + lvl_words, *dpts = nxt + lvl_words;
    MS_CUDA(c, cudaMemcpyAsync(dpts, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice, c->stream));
    {   // level 0: the coefficient matrix itself, all points at once
        dim3 grid((unsigned)nb0, ncols, K);
        if (field == 1)
            poly_eval_kernel<1><<<grid, kPeThreads, 0, c->stream>>>(in.as<u64>(), col_stride_elems, n, dpts, cur, (size_t)K * nb0);
        else
            poly_eval_kernel<3><<<grid, kPeThreads, 0, c->stream>>>(in.as<u64>(), col_stride_elems * 3, n, dpts, cur, (size_t)K * nb0);
        c->launches++;
        MS_CHECK_LAUNCH(c);
    }
    size_t count = nb0;
    while (count > 1) {   // partials [(col*K + k)*count + b] are Fq3 coefficients of a polynomial in x^span
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (unsigned k = 0; k < K; k++) {
            Fq3 q = gl::pow(Fq3{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]}, (u64)kPeSpan);
            pts[3 * k] = q.c0; pts[3 * k + 1] = q.c1; pts[3 * k + 2] = q.c2;
        }
        MS_CUDA(c, cudaMemcpyAsync(dpts, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice, c->stream));
        const size_t nb = (count + kPeSpan - 1) / kPeSpan;
        for (unsigned k = 0; k < K; k++) {
            dim3 grid((unsigned)nb, ncols, 1);
            poly_eval_kernel<3><<<grid, kPeThreads, 0, c->stream>>>(cur + 3 * (size_t)k * count, 3 * (size_t)K * count, count,
                                                                    dpts + 3 * k, nxt + 3 * (size_t)k * nb, (size_t)K * nb);
            c->launches++;
        }
        MS_CHECK_LAUNCH(c);
        u64 *t = cur; cur = nxt; nxt = t;
        count = nb;
    }
    MS_CUDA(c, cudaMemcpyAsync(out, cur, 3 * (size_t)ncols * K * 8, cudaMemcpyDefault, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return in.finish();
}
