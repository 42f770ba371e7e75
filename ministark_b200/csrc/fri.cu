// fri.cu — FRI degree-respecting projection, evaluated coset by coset.
//
// The reference computes the next codeword as
//     bit_reverse -> iNTT(n) -> *ff -> fold by alpha -> NTT(n/ff) -> bit_reverse
// (apply_drp, src/fri.rs:526-567: two full-size transforms and two CPU permutations per
// layer).  Because the codeword is kept in bit-reversed order, the ff evaluations of the
// coset  x_k * <omega_ff>,  x_k = offset * g^bitrev(k),  sit next to each other (row k of the
// committed layer matrix, src/fri.rs:199-231), and
//     out[k] = ff * P_k(alpha),   P_k = the interpolant of row k over its coset
// which is what the verifier recomputes (src/fri.rs:393-412).  One thread per row:
//     D_j = sum_i v_i omega_ff^(-ij)      (in-register inverse-direction DFT, size ff <= 16)
//     out[k] = sum_j D_j (alpha / x_k)^j  (Horner)
// One pass over the codeword: s*M bytes read, s*M/ff written.
#include "ctx.cuh"
#include "dft.cuh"

namespace ms {

using gl::Fq3;
using msntt::brev_c;
using msntt::IC;
using msntt::static_for;

// rows are FF*LANES consecutive words; a CTA first copies its block of rows into shared memory with fully
// coalesced loads (row pitch padded by one word: conflict-free row reads), then each thread folds one row
template <int LOGFF, int LANES>
__global__ void __launch_bounds__(128) fri_fold_kernel(const u64 *__restrict__ evals, u64 *__restrict__ out, size_t m,
                                                        unsigned log_m, u64 offset_inv, const u64 *__restrict__ tw_lo, const u64 *__restrict__ tw_hi,
                                                        unsigned hi_len, u64 n_mask, u64 a0, u64 a1, u64 a2) {
    constexpr int FF = 1 << LOGFF;
    constexpr int RW = FF * LANES, PITCH = RW + 1;
    extern __shared__ u64 rows_sm[];
    const size_t k0 = blockIdx.x * (size_t)blockDim.x;
    const size_t k = k0 + threadIdx.x;
    {
        const size_t rows_here = (m - k0) < (size_t)blockDim.x ? (m - k0) : (size_t)blockDim.x;
        const size_t words = rows_here * RW;
        const u64 *blk = evals + k0 * RW;
        for (size_t w = threadIdx.x; w < words; w += blockDim.x) rows_sm[(w / RW) * PITCH + (w % RW)] = blk[w];
    }
    __syncthreads();
    if (k >= m) return;
    const u64 *src = rows_sm + (size_t)threadIdx.x * PITCH;
    u64 D[LANES][FF];
#pragma unroll
    for (int l = 0; l < LANES; l++) {
        u64 x[FF];
        // entry t of the row is the evaluation with natural index bitrev(t) inside the coset
        static_for<0, FF>([&](auto T) { x[brev_c(decltype(T)::value, LOGFF)] = src[decltype(T)::value * LANES + l]; });
        msntt::dft_regs<LOGFF, true>(x);
        static_for<0, FF>([&](auto J) { D[l][decltype(J)::value] = gl::canon(x[brev_c(decltype(J)::value, LOGFF)]); });
    }
    // 1 / x_k = offset^-1 * g^-bitrev_m(k) = offset^-1 * g^(n - bitrev_m(k)), from the two-level table of g_n^e
    const u64 e = log_m ? (__brevll((u64)k) >> (64 - log_m)) : 0;
    const u64 ne = (n_mask + 1 - e) & n_mask;
    u64 gpow = tw_lo[ne & 4095];
    if (hi_len > 1) gpow = gl::mul(tw_hi[ne >> 12], gpow);
    const u64 xinv = gl::mul(offset_inv, gpow);
    if constexpr (LANES == 1) {
        const u64 beta = gl::mul(a0, xinv);
        u64 acc = D[0][FF - 1];
#pragma unroll
        for (int j = FF - 2; j >= 0; j--) acc = gl::add(gl::mul(acc, beta), D[0][j]);
        out[k] = acc;
    } else {
        const Fq3 beta = gl::mul(Fq3{a0, a1, a2}, xinv);
        Fq3 acc{D[0][FF - 1], D[1][FF - 1], D[2][FF - 1]};
#pragma unroll
        for (int j = FF - 2; j >= 0; j--) acc = gl::add(gl::mul(acc, beta), Fq3{D[0][j], D[1][j], D[2][j]});
        out[3 * k] = acc.c0;
        out[3 * k + 1] = acc.c1;
        out[3 * k + 2] = acc.c2;
    }
}

}  // namespace ms

using namespace ms;

extern "C" int ms_fri_fold(ms_ctx *c, int field, const void *evals, unsigned log_n, unsigned log_ff, uint64_t offset_mont,
                           const uint64_t *alpha, void *out) {
    if (!c || !evals || !alpha || !out) return MS_ERR_INVALID;
    if (field != 1 && field != 3) return fail(c, MS_ERR_INVALID, "ms_fri_fold: bad field id");
    // folding factors the reference supports: 2, 4, 8, 16 (src/fri.rs:185-192)
    if (log_ff < 1 || log_ff > 4 || log_ff > log_n || log_n > 32) return fail(c, MS_ERR_INVALID, "ms_fri_fold: bad sizes");
    if (offset_mont >= gl::P || offset_mont == 0) return fail(c, MS_ERR_INVALID, "offset must be a non-zero canonical word");
    const size_t n = (size_t)1 << log_n, m = n >> log_ff;
    u64 a[3] = {0, 0, 0};
    if (is_device_ptr(alpha)) MS_CUDA(c, cudaMemcpy(a, alpha, field * 8, cudaMemcpyDeviceToHost));
    else for (int i = 0; i < field; i++) a[i] = alpha[i];
    Staged in(c, evals, n * field * 8, true, false);
    if (in.rc) return in.rc;
    Staged o(c, out, m * field * 8, false, true);
    if (o.rc) return o.rc;
    const u64 off_inv = gl::inv(offset_mont);
    const u64 *tw_lo, *tw_hi;
    u32 hi_len;
    if (int trc = ntt_plan_tables(c, log_n, &tw_lo, &tw_hi, &hi_len)) return trc;
    const u64 n_mask = (u64)n - 1;
    const unsigned threads = (field == 3 && log_ff == 4) ? 64 : 128, blocks = (unsigned)((m + threads - 1) / threads),
                   log_m = log_n - log_ff;
    const size_t smem = (size_t)threads * (((size_t)field << log_ff) + 1) * 8;
#define MS_FF(L, F) fri_fold_kernel<L, F><<<blocks, threads, smem, c->stream>>>(in.as<u64>(), o.as<u64>(), m, log_m, off_inv, tw_lo, tw_hi, hi_len, n_mask, a[0], a[1], a[2])
    if (field == 1) {
        switch (log_ff) { case 1: MS_FF(1, 1); break; case 2: MS_FF(2, 1); break; case 3: MS_FF(3, 1); break; default: MS_FF(4, 1); }
    } else {
        switch (log_ff) { case 1: MS_FF(1, 3); break; case 2: MS_FF(2, 3); break; case 3: MS_FF(3, 3); break; default: MS_FF(4, 3); }
    }
#undef MS_FF
    c->launches++;
    MS_CHECK_LAUNCH(c);
    int rc;
    if ((rc = in.finish())) return rc;
    return o.finish();
}
