/*
 * gl_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (andrewmilson/ministark @ ff5bf80)
 * parallel CPU prover path for the gpu-poly hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (ministark_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no numeric golden vectors for this path
 *  and its Rust toolchain / arkworks crates are not available in
 * the build container, so this oracle cannot be checked against reference
 * output.  It is pinned instead by (tests/test_oracle_*.py):
 *   - the constants the reference does carry (ONE, R2, NONRESIDUE, bit_reverse
 *     golden permutation, gpu/src/utils.rs:233-236),
 *   - an independent big-integer Python spec (oracle/pyspec.py) incl. O(n^2) DFT,
 *   - hashlib SHA-256.
 *
 * Third-party algorithm sources restated here (absent from /root/reference):
 *   ark-ff 0.4.2 / ark-ff-optimized 0.4.1 (Goldilocks fp64, Montgomery R=2^64),
 *   ark-poly 0.4.2 Radix2EvaluationDomain (in-order FFT = DIF + derange,
 *   in-order iFFT = derange + DIT, degree-aware FFT), sha2 0.10.8 (FIPS 180-4).
 *
 * Conventions: p = 2^64 - 2^32 + 1.  Every u64 "word" is a canonical (< p)
 * Montgomery residue x*2^64 mod p (gpu/src/metal/felt_u64.h.metal:118,127).
 * Fq3 = Fp[X]/(X^3-2), stored [c0,c1,c2] (gpu/src/fields.rs:78-97).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_NEG_PINV 0xFFFFFFFEFFFFFFFFULL /* -p^{-1} mod 2^64 ; p^{-1} = 2^32+1 */
#define GL_ONE 0xFFFFFFFFULL              /* 2^64 mod p (felt_u64.h.metal:118) */
#define GL_R2 0xFFFFFFFE00000001ULL       /* 2^128 mod p (felt_u64.h.metal:127) */
/* canonical 2^32-th root of unity 7^((p-1)/2^32); generator 7 (src/air.rs:42-44) */
#define GL_TWO_ADIC_ROOT_CANON 1753635133440165772ULL
#define GL_TWO_ADICITY 32

/* ---------------------------------------------------------------- Fp ---- */
/* semantics of felt_u64.h.metal:147-177 (add / sub / Montgomery mul) */
static inline u64 fp_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
static inline u64 fp_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
static inline u64 fp_neg(u64 a) { return a ? GL_P - a : 0; }
static inline u64 fp_mul(u64 a, u64 b) {
    u128 x = (u128)a * b;
    u64 m = (u64)x * GL_NEG_PINV;
    u128 mp = (u128)m * GL_P;
    u64 lo_sum = (u64)x + (u64)mp; /* == 0 by construction */
    u64 carry = lo_sum < (u64)x;
    u128 t = (x >> 64) + (mp >> 64) + carry;
    if (t >= GL_P) t -= GL_P;
    return (u64)t;
}
static inline u64 fp_from_canon(u64 x) { return fp_mul(x % GL_P, GL_R2); }
static inline u64 fp_to_canon(u64 x) { return fp_mul(x, 1); }
static u64 fp_pow(u64 a, u64 e) {
    u64 r = GL_ONE;
    while (e) {
        if (e & 1) r = fp_mul(r, a);
        a = fp_mul(a, a);
        e >>= 1;
    }
    return r;
}
static inline u64 fp_inv(u64 a) { return fp_pow(a, GL_P - 2); }

/* --------------------------------------------------------------- Fq3 ---- */
/* X^3 = 2.  Schoolbook product (same field element as the reference's
 * Karatsuba form, felt_u64.h.metal:205-231). */
typedef struct { u64 c[3]; } fq3;
static inline fq3 fq3_add(fq3 a, fq3 b) {
    fq3 r = {{fp_add(a.c[0], b.c[0]), fp_add(a.c[1], b.c[1]), fp_add(a.c[2], b.c[2])}};
    return r;
}
static inline fq3 fq3_sub(fq3 a, fq3 b) {
    fq3 r = {{fp_sub(a.c[0], b.c[0]), fp_sub(a.c[1], b.c[1]), fp_sub(a.c[2], b.c[2])}};
    return r;
}
static inline fq3 fq3_mul(fq3 a, fq3 b) {
    u64 a0b0 = fp_mul(a.c[0], b.c[0]), a0b1 = fp_mul(a.c[0], b.c[1]), a0b2 = fp_mul(a.c[0], b.c[2]);
    u64 a1b0 = fp_mul(a.c[1], b.c[0]), a1b1 = fp_mul(a.c[1], b.c[1]), a1b2 = fp_mul(a.c[1], b.c[2]);
    u64 a2b0 = fp_mul(a.c[2], b.c[0]), a2b1 = fp_mul(a.c[2], b.c[1]), a2b2 = fp_mul(a.c[2], b.c[2]);
    u64 t3 = fp_add(a1b2, a2b1); /* X^3 coefficient */
    u64 t4 = a2b2;               /* X^4 coefficient */
    fq3 r;
    r.c[0] = fp_add(a0b0, fp_add(t3, t3));
    r.c[1] = fp_add(fp_add(a0b1, a1b0), fp_add(t4, t4));
    r.c[2] = fp_add(fp_add(a0b2, a1b1), a2b0);
    return r;
}
static inline fq3 fq3_mul_fp(fq3 a, u64 b) {
    fq3 r = {{fp_mul(a.c[0], b), fp_mul(a.c[1], b), fp_mul(a.c[2], b)}};
    return r;
}
static inline fq3 fq3_from_fp(u64 a) { fq3 r = {{a, 0, 0}}; return r; }
static inline fq3 fq3_one(void) { return fq3_from_fp(GL_ONE); }
static inline fq3 fq3_zero(void) { return fq3_from_fp(0); }
static fq3 fq3_pow(fq3 a, u64 e) {
    fq3 r = fq3_one();
    while (e) {
        if (e & 1) r = fq3_mul(r, a);
        a = fq3_mul(a, a);
        e >>= 1;
    }
    return r;
}
/* inverse via the adjugate of the multiplication matrix over X^3 - 2
 * (ark-ff CubicExtField::inverse, "Algorithm 17" of ePrint 2010/354) */
static fq3 fq3_inv(fq3 a) {
    u64 two = fp_add(GL_ONE, GL_ONE);
    u64 t0 = fp_mul(a.c[0], a.c[0]), t1 = fp_mul(a.c[1], a.c[1]), t2 = fp_mul(a.c[2], a.c[2]);
    u64 t3 = fp_mul(a.c[0], a.c[1]), t4 = fp_mul(a.c[0], a.c[2]), t5 = fp_mul(a.c[1], a.c[2]);
    u64 s0 = fp_sub(t0, fp_mul(two, t5));
    u64 s1 = fp_sub(fp_mul(two, t2), t3);
    u64 s2 = fp_sub(t1, t4);
    u64 a1 = fp_mul(a.c[2], s1), a2 = fp_mul(a.c[1], s2);
    u64 d = fp_add(fp_mul(a.c[0], s0), fp_mul(two, fp_add(a1, a2)));
    u64 di = fp_inv(d);
    fq3 r = {{fp_mul(s0, di), fp_mul(s1, di), fp_mul(s2, di)}};
    return r;
}

/* ------------------------------------------------------ exported field -- */
u64 orc_fp_one(void) { return GL_ONE; }
u64 orc_fp_from_canonical(u64 x) { return fp_from_canon(x); }
u64 orc_fp_to_canonical(u64 x) { return fp_to_canon(x); }
u64 orc_fp_mul1(u64 a, u64 b) { return fp_mul(a, b); }
u64 orc_fp_add1(u64 a, u64 b) { return fp_add(a, b); }
u64 orc_fp_sub1(u64 a, u64 b) { return fp_sub(a, b); }
u64 orc_fp_inv1(u64 a) { return fp_inv(a); }
u64 orc_fp_pow1(u64 a, u64 e) { return fp_pow(a, e); }
void orc_vec_from_canonical(u64 *v, size_t n) { for (size_t i = 0; i < n; i++) v[i] = fp_from_canon(v[i]); }
void orc_vec_to_canonical(u64 *v, size_t n) { for (size_t i = 0; i < n; i++) v[i] = fp_to_canon(v[i]); }
void orc_fq3_mul1(const u64 *a, const u64 *b, u64 *out) {
    fq3 x, y; memcpy(&x, a, 24); memcpy(&y, b, 24);
    fq3 r = fq3_mul(x, y); memcpy(out, &r, 24);
}
void orc_fq3_inv1(const u64 *a, u64 *out) { fq3 x; memcpy(&x, a, 24); fq3 r = fq3_inv(x); memcpy(out, &r, 24); }
void orc_fq3_pow1(const u64 *a, u64 e, u64 *out) { fq3 x; memcpy(&x, a, 24); fq3 r = fq3_pow(x, e); memcpy(out, &r, 24); }

/* root of unity of order 2^log_n, Montgomery form
 * (ark-ff FftField::get_root_of_unity: TWO_ADIC_ROOT^(2^(32-log_n))) */
u64 orc_root_of_unity(unsigned log_n) {
    u64 r = fp_from_canon(GL_TWO_ADIC_ROOT_CANON);
    for (unsigned i = log_n; i < GL_TWO_ADICITY; i++) r = fp_mul(r, r);
    return r;
}
u64 orc_generator(void) { return fp_from_canon(7); }

/* --------------------------------------------------------- bit reverse -- */
static inline size_t bitrev(size_t i, unsigned log_n) {
    if (log_n == 0) return 0;
    u64 x = (u64)i;
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = __builtin_bswap64(x);
    return (size_t)(x >> (64 - log_n));
}
/* gpu/src/utils.rs:4-11,32-78.  elem_words = 1 (Fp) or 3 (Fq3).  Like the reference's parallel build: split over the
 * threads from 2^17 elements up (every pair is swapped by its smaller index, so the iterations are independent);
 * inside an enclosing parallel region (one column per thread) the loop stays serial. */
void orc_bit_reverse(u64 *v, unsigned elem_words, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    #pragma omp parallel for if (log_n >= 17) schedule(static)
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, log_n);
        if (j > i)
            for (unsigned w = 0; w < elem_words; w++) {
                u64 t = v[i * elem_words + w];
                v[i * elem_words + w] = v[j * elem_words + w];
                v[j * elem_words + w] = t;
            }
    }
}

/* ----------------------------------------------------------------- NTT -- */
/* One column with `lanes` interleaved Fp lanes (1 = Fp, 3 = Fq3; an Fq3 NTT
 * uses Fp twiddles, fft_shaders.h.metal:152-163).  In-place radix-2. */

static u64 *roots_table(u64 root, size_t half) {
    u64 *r = (u64 *)malloc(sizeof(u64) * (half ? half : 1));
    u64 acc = GL_ONE;
    for (size_t i = 0; i < half; i++) { r[i] = acc; acc = fp_mul(acc, root); }
    return r;
}

/* DIF: natural in -> bit-reversed out (ark-poly io_helper) */
static void dif_levels(u64 *a, unsigned lanes, unsigned log_n, const u64 *roots, int par) {
    size_t n = (size_t)1 << log_n;
    for (size_t gap = n >> 1; gap >= 1; gap >>= 1) {
        const size_t step = (n >> 1) / gap, nblk = n / (2 * gap);
        if (nblk >= 64 || !par) {
            #pragma omp parallel for if (par) schedule(static)
            for (size_t blk = 0; blk < nblk; blk++) {
                u64 *lo = a + blk * 2 * gap * lanes, *hi = lo + gap * lanes;
                for (size_t j = 0; j < gap; j++) {
                    const u64 w = roots[j * step];
                    for (unsigned l = 0; l < lanes; l++) {
                        const u64 u = lo[j * lanes + l], v = hi[j * lanes + l];
                        lo[j * lanes + l] = fp_add(u, v);
                        hi[j * lanes + l] = fp_mul(fp_sub(u, v), w);
                    }
                }
            }
        } else {
            for (size_t blk = 0; blk < nblk; blk++) {
                u64 *lo = a + blk * 2 * gap * lanes, *hi = lo + gap * lanes;
                #pragma omp parallel for schedule(static)
                for (size_t j = 0; j < gap; j++) {
                    const u64 w = roots[j * step];
                    for (unsigned l = 0; l < lanes; l++) {
                        const u64 u = lo[j * lanes + l], v = hi[j * lanes + l];
                        lo[j * lanes + l] = fp_add(u, v);
                        hi[j * lanes + l] = fp_mul(fp_sub(u, v), w);
                    }
                }
            }
        }
    }
}
/* DIT: bit-reversed in -> natural out, starting at `start_gap`
 * (ark-poly oi_helper(x, root, start_gap)) */
static void dit_levels(u64 *a, unsigned lanes, unsigned log_n, const u64 *roots, size_t start_gap, int par) {
    size_t n = (size_t)1 << log_n;
    for (size_t gap = start_gap; gap < n; gap <<= 1) {
        const size_t step = (n >> 1) / gap, nblk = n / (2 * gap);
        if (nblk >= 64 || !par) {
            #pragma omp parallel for if (par) schedule(static)
            for (size_t blk = 0; blk < nblk; blk++) {
                u64 *lo = a + blk * 2 * gap * lanes, *hi = lo + gap * lanes;
                for (size_t j = 0; j < gap; j++) {
                    const u64 w = roots[j * step];
                    for (unsigned l = 0; l < lanes; l++) {
                        const u64 u = lo[j * lanes + l], v = fp_mul(hi[j * lanes + l], w);
                        lo[j * lanes + l] = fp_add(u, v);
                        hi[j * lanes + l] = fp_sub(u, v);
                    }
                }
            }
        } else {
            for (size_t blk = 0; blk < nblk; blk++) {
                u64 *lo = a + blk * 2 * gap * lanes, *hi = lo + gap * lanes;
                #pragma omp parallel for schedule(static)
                for (size_t j = 0; j < gap; j++) {
                    const u64 w = roots[j * step];
                    for (unsigned l = 0; l < lanes; l++) {
                        const u64 u = lo[j * lanes + l], v = fp_mul(hi[j * lanes + l], w);
                        lo[j * lanes + l] = fp_add(u, v);
                        hi[j * lanes + l] = fp_sub(u, v);
                    }
                }
            }
        }
    }
}
/* a[i] *= c * g^i  (ark-poly distribute_powers_and_mul_by_const) */
static void distribute_powers(u64 *a, unsigned lanes, size_t n, u64 g, u64 c, int par) {
    const size_t chunk = 1024;
    size_t nch = (n + chunk - 1) / chunk;
    #pragma omp parallel for if (par) schedule(static)
    for (size_t ch = 0; ch < nch; ch++) {
        size_t s = ch * chunk, e = s + chunk < n ? s + chunk : n;
        u64 acc = fp_mul(c, fp_pow(g, s));
        for (size_t i = s; i < e; i++) {
            for (unsigned l = 0; l < lanes; l++) a[i * lanes + l] = fp_mul(a[i * lanes + l], acc);
            acc = fp_mul(acc, g);
        }
    }
}

/* Work split of a batch of independent columns (the reference: one rayon task per column, src/matrix.rs:118-139, with
 * ark-poly's own parallel FFT inside a task).  Whole columns per thread, unless that would leave half of the threads or
 * more without a column: then the columns run one after the other with every butterfly level split over all threads
 * (measured on 8 threads, 2^20 -> 2^23 LDE: 2 columns 0.92 s instead of 2.28 s, 4 columns 1.83 s instead of 2.58 s; from
 * 8 columns up whole columns win).  So 32 columns on 32 threads (config 3) or 17 on 8 run by columns; the 8 columns of
 * examples/fib on a 32-thread host run level-parallel.  ORACLE_COLUMN_PARALLEL=0|1 forces one or the other (A/B). */
static int column_parallel(unsigned ncols) {
#ifdef _OPENMP
    const unsigned t = (unsigned)omp_get_max_threads();
#else
    const unsigned t = 1;
#endif
    const char *e = getenv("ORACLE_COLUMN_PARALLEL");
    if (ncols <= 1 || t <= 1) return 0;
    if (e && *e) return *e != '0';
    return 2 * ncols > t;
}

static void ntt_column(u64 *a, unsigned lanes, unsigned log_n, u64 offset, int inverse,
                       const u64 *roots, int par) {
    size_t n = (size_t)1 << log_n;
    if (log_n < 12) par = 0;   /* a few thousand butterflies: forking a team per level costs more than the level */
    if (!inverse) {
        /* in_order_fft_in_place: distribute offset powers, DIF, derange */
        if (offset != GL_ONE) distribute_powers(a, lanes, n, offset, GL_ONE, par);
        dif_levels(a, lanes, log_n, roots, par);
        orc_bit_reverse(a, lanes, log_n);
    } else {
        /* in_order_ifft_in_place: derange, DIT with inverse root, scale */
        orc_bit_reverse(a, lanes, log_n);
        dit_levels(a, lanes, log_n, roots, 1, par);
        u64 n_inv = fp_inv(fp_from_canon((u64)n));
        u64 off_inv = fp_inv(offset);
        distribute_powers(a, lanes, n, off_inv, n_inv, par);
    }
}

/* Column-major matrix: column c starts at base + c*col_stride_words.
 * Natural order in, natural order out.  Mirrors Matrix::into_polynomials_cpu /
 * into_evaluations_cpu (src/matrix.rs:118-139,165-190): one task per column. */
void orc_ntt_columns(u64 *base, size_t col_stride_words, unsigned ncols, unsigned lanes,
                     unsigned log_n, u64 offset_mont, int inverse) {
    size_t n = (size_t)1 << log_n;
    u64 root = orc_root_of_unity(log_n);
    if (inverse) root = fp_inv(root);
    u64 *roots = roots_table(root, n >> 1);
    int par_cols = column_parallel(ncols);
    #pragma omp parallel for if (par_cols) schedule(dynamic, 1)
    for (unsigned c = 0; c < ncols; c++)
        ntt_column(base + (size_t)c * col_stride_words, lanes, log_n, offset_mont, inverse, roots, !par_cols);
    free(roots);
}

/* Low-degree extension: n coefficients -> N = n << log_blowup evaluations over
 * the coset offset*<g_N>, natural order, or bit-reversed order when bitrev != 0
 * (Matrix::into_evaluations / into_bit_reversed_evaluations, src/matrix.rs:165-234).
 * Follows ark-poly's degree-aware FFT for blowup >= 4 (zero-padded DIF+derange
 * otherwise): scale by offset^i, place coefficient i at bitrev_N(i), duplicate,
 * DIT from gap = blowup. */
void orc_lde_columns(const u64 *in_base, size_t in_stride_words, u64 *out_base, size_t out_stride_words,
                     unsigned ncols, unsigned lanes, unsigned log_n, unsigned log_blowup,
                     u64 offset_mont, int bitrev_out) {
    unsigned log_N = log_n + log_blowup;
    size_t n = (size_t)1 << log_n, N = (size_t)1 << log_N, dup = (size_t)1 << log_blowup;
    u64 root = orc_root_of_unity(log_N);
    u64 *roots = roots_table(root, N >> 1);
    int par_cols = column_parallel(ncols);
    const int par_in = !par_cols && log_N >= 12;
    #pragma omp parallel for if (par_cols) schedule(dynamic, 1)
    for (unsigned c = 0; c < ncols; c++) {
        const u64 *in = in_base + (size_t)c * in_stride_words;
        u64 *out = out_base + (size_t)c * out_stride_words;
        if (log_blowup >= 2) {
            u64 *tmp = (u64 *)malloc(sizeof(u64) * n * lanes);
            memcpy(tmp, in, sizeof(u64) * n * lanes);
            if (offset_mont != GL_ONE) distribute_powers(tmp, lanes, n, offset_mont, GL_ONE, par_in);
            #pragma omp parallel for if (par_in) schedule(static)
            for (size_t i = 0; i < n; i++) {
                size_t ri = bitrev(i, log_N); /* multiple of dup */
                for (size_t d = 0; d < dup; d++)
                    for (unsigned l = 0; l < lanes; l++) out[(ri + d) * lanes + l] = tmp[i * lanes + l];
            }
            free(tmp);
            dit_levels(out, lanes, log_N, roots, dup, par_in);
        } else {
            memset(out, 0, sizeof(u64) * N * lanes);
            memcpy(out, in, sizeof(u64) * n * lanes);
            if (offset_mont != GL_ONE) distribute_powers(out, lanes, n, offset_mont, GL_ONE, par_in);
            dif_levels(out, lanes, log_N, roots, par_in);
            orc_bit_reverse(out, lanes, log_N);
        }
        if (bitrev_out) orc_bit_reverse(out, lanes, log_N);
    }
    free(roots);
}

/* ------------------------------------------------------------- SHA-256 -- */
static const u32 K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline u32 rotr(u32 x, int r) { return (x >> r) | (x << (32 - r)); }

/* The reference's sha2 0.10.8 selects the SHA-NI backend at run time on x86-64 (cpufeatures); for a fair
 * CPU baseline the oracle does the same.  Standard SHA extensions schedule: state kept as ABEF / CDGH. */
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
static int have_sha_ni(void) {
    static int cached = -1;
    if (cached < 0) {
        unsigned a, b, c, d;
        cached = (__get_cpuid_count(7, 0, &a, &b, &c, &d) && (b & (1u << 29))) ? 1 : 0;
        if (getenv("ORACLE_NO_SHA_NI")) cached = 0;
    }
    return cached;
}
__attribute__((target("sha,sse4.1,ssse3")))
static void sha256_block_ni(u32 st[8], const uint8_t *blk) {
    const __m128i mask = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&st[0]);
    __m128i s1 = _mm_loadu_si128((const __m128i *)&st[4]);
    tmp = _mm_shuffle_epi32(tmp, 0xB1);
    s1 = _mm_shuffle_epi32(s1, 0x1B);
    __m128i s0 = _mm_alignr_epi8(tmp, s1, 8);
    s1 = _mm_blend_epi16(s1, tmp, 0xF0);
    const __m128i save0 = s0, save1 = s1;
    __m128i m[4];
    for (int g = 0; g < 16; g++) {
        __m128i w;
        if (g < 4) {
            w = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(blk + 16 * g)), mask);
        } else {
            __m128i t = _mm_add_epi32(_mm_sha256msg1_epu32(m[0], m[1]), _mm_alignr_epi8(m[3], m[2], 4));
            w = _mm_sha256msg2_epu32(t, m[3]);
            m[0] = m[1]; m[1] = m[2]; m[2] = m[3];
        }
        m[g < 4 ? g : 3] = w;
        __m128i msg = _mm_add_epi32(w, _mm_loadu_si128((const __m128i *)&K256[4 * g]));
        s1 = _mm_sha256rnds2_epu32(s1, s0, msg);
        msg = _mm_shuffle_epi32(msg, 0x0E);
        s0 = _mm_sha256rnds2_epu32(s0, s1, msg);
    }
    s0 = _mm_add_epi32(s0, save0);
    s1 = _mm_add_epi32(s1, save1);
    tmp = _mm_shuffle_epi32(s0, 0x1B);
    s1 = _mm_shuffle_epi32(s1, 0xB1);
    s0 = _mm_blend_epi16(tmp, s1, 0xF0);
    s1 = _mm_alignr_epi8(s1, tmp, 8);
    _mm_storeu_si128((__m128i *)&st[0], s0);
    _mm_storeu_si128((__m128i *)&st[4], s1);
}
#else
static int have_sha_ni(void) { return 0; }
static void sha256_block_ni(u32 st[8], const uint8_t *blk) { (void)st; (void)blk; }
#endif
int orc_have_sha_ni(void) { return have_sha_ni(); }

static void sha256_block_portable(u32 st[8], const uint8_t *blk);
static inline void sha256_block(u32 st[8], const uint8_t *blk) {
    if (have_sha_ni()) sha256_block_ni(st, blk);
    else sha256_block_portable(st, blk);
}
static void sha256_block_portable(u32 st[8], const uint8_t *blk) {
    u32 w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((u32)blk[4 * i] << 24) | ((u32)blk[4 * i + 1] << 16) | ((u32)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        u32 s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        u32 s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        u32 S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        u32 ch = (e & f) ^ (~e & g);
        u32 t1 = h + S1 + ch + K256[i] + w[i];
        u32 S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        u32 mj = (a & b) ^ (a & c) ^ (b & c);
        u32 t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
void orc_sha256(const uint8_t *msg, size_t len, uint8_t out[32]) {
    u32 st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t off = 0;
    for (; off + 64 <= len; off += 64) sha256_block(st, msg + off);
    uint8_t tail[128];
    size_t rem = len - off;
    memset(tail, 0, sizeof tail);
    memcpy(tail, msg + off, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 9 <= 64 ? 64 : 128;
    u64 bits = (u64)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}

/* leaf_i = SHA-256( || over columns c, lanes l of LE64(canonical(col_c[i].l)) )
 * (src/hash.rs:92-99 + ark-serialize: Fp -> into_bigint() 8-byte LE,
 *  Fq3 -> c0||c1||c2; src/merkle.rs:412-436 row chunks across threads) */
void orc_hash_rows(const u64 *base, size_t col_stride_words, unsigned ncols, unsigned lanes,
                   size_t nrows, uint8_t *digests) {
    size_t row_bytes = (size_t)ncols * lanes * 8;
    #pragma omp parallel
    {
        uint8_t *buf = (uint8_t *)malloc(row_bytes ? row_bytes : 1);
        #pragma omp for schedule(static)
        for (size_t i = 0; i < nrows; i++) {
            uint8_t *q = buf;
            for (unsigned c = 0; c < ncols; c++)
                for (unsigned l = 0; l < lanes; l++) {
                    u64 x = fp_to_canon(base[(size_t)c * col_stride_words + i * lanes + l]);
                    for (int b = 0; b < 8; b++) *q++ = (uint8_t)(x >> (8 * b));
                }
            orc_sha256(buf, row_bytes, digests + 32 * i);
        }
        free(buf);
    }
}

/* nodes[k] = SHA-256(nodes[2k] || nodes[2k+1]); level of leaf pairs at
 * nodes[n/2 + i] = SHA-256(leaf[2i] || leaf[2i+1]); nodes[0] = 32 zero bytes
 * (Digest::default); root = nodes[1]  (src/merkle.rs:438-508, src/hash.rs:77-82) */
void orc_merkle_nodes(const uint8_t *leaves, size_t n, uint8_t *nodes) {
    memset(nodes, 0, 32);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n / 2; i++) orc_sha256(leaves + 64 * i, 64, nodes + 32 * (n / 2 + i));
    for (size_t size = n / 4; size >= 1; size >>= 1) {
        #pragma omp parallel for schedule(static) if (size >= 1024)
        for (size_t k = size; k < 2 * size; k++) orc_sha256(nodes + 64 * k, 64, nodes + 32 * k);
    }
}

/* ------------------------------------------------------- pointwise ops -- */
/* Element-wise stage semantics of gpu/src/metal/evaluation_shaders.h.metal:11-168.
 * op codes shared with include/ministark_b200.h (MS_OP_*).
 * lfield/rfield: 1 = Fp, 3 = Fq3.  rhs index is (i + shift) % n. */
enum { OP_MUL = 0, OP_ADD = 1, OP_CONVERT = 2, OP_INV = 3, OP_EXP = 4, OP_NEG = 5, OP_MULPOW = 6, OP_FILL = 7, OP_SUB = 8 };

static fq3 load_el(const u64 *p, unsigned f, size_t i) {
    fq3 r;
    if (f == 1) { r.c[0] = p[i]; r.c[1] = 0; r.c[2] = 0; }
    else memcpy(&r, p + 3 * i, 24);
    return r;
}
static void store_el(u64 *p, unsigned f, size_t i, fq3 v) {
    if (f == 1) p[i] = v.c[0]; else memcpy(p + 3 * i, &v, 24);
}
/* batch inversion over chunks of 512 elements, as eval_cpu.rs:280-295 / ark_ff::batch_inversion does:
 * one field inversion per chunk + 3 multiplications per element (zeros are skipped and stay zero) */
static void batch_inverse(unsigned field, u64 *dst, const u64 *src, size_t n) {
    const size_t CH = 512;
    size_t nch = (n + CH - 1) / CH;
    #pragma omp parallel for schedule(static)
    for (size_t ch = 0; ch < nch; ch++) {
        size_t s0 = ch * CH, e0 = s0 + CH < n ? s0 + CH : n;
        fq3 pref[512];
        fq3 acc = fq3_one();
        for (size_t i = s0; i < e0; i++) {
            fq3 v = load_el(src, field, i);
            int z = (v.c[0] | v.c[1] | v.c[2]) == 0;
            pref[i - s0] = acc;
            if (!z) acc = fq3_mul(acc, v);
        }
        fq3 inv = (field == 1) ? fq3_from_fp(fp_inv(acc.c[0])) : fq3_inv(acc);
        for (size_t i = e0; i-- > s0;) {
            fq3 v = load_el(src, field, i);
            int z = (v.c[0] | v.c[1] | v.c[2]) == 0;
            if (z) { store_el(dst, field, i, fq3_zero()); continue; }
            store_el(dst, field, i, fq3_mul(inv, pref[i - s0]));
            inv = fq3_mul(inv, v);
        }
    }
}

/* dst[i] = lhs[i] OP rhs[(i+shift)%n]   (MulInto/AddInto; *Assign when dst==lhs) */
void orc_pointwise(int op, unsigned dfield, u64 *dst, unsigned lfield, const u64 *lhs,
                   unsigned rfield, const u64 *rhs, size_t n, size_t shift, u64 exponent) {
    if (op == OP_INV && dfield == lfield) {
        if (dst == lhs) {
            u64 *tmp = (u64 *)malloc(sizeof(u64) * n * lfield);
            memcpy(tmp, lhs, sizeof(u64) * n * lfield);
            batch_inverse(lfield, dst, tmp, n);
            free(tmp);
        } else {
            batch_inverse(lfield, dst, lhs, n);
        }
        return;
    }
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        fq3 a = load_el(lhs, lfield, i), r;
        fq3 b = rhs ? load_el(rhs, rfield, (i + shift) % n) : fq3_zero();
        switch (op) {
        case OP_MUL: r = fq3_mul(a, b); break;
        case OP_ADD: r = fq3_add(a, b); break;
        case OP_SUB: r = fq3_sub(a, b); break;
        case OP_CONVERT: r = a; break;
        case OP_INV: r = (lfield == 1) ? fq3_from_fp(fp_inv(a.c[0])) : fq3_inv(a); break;
        case OP_EXP: r = fq3_pow(a, exponent); break;
        case OP_NEG: r.c[0] = fp_neg(a.c[0]); r.c[1] = fp_neg(a.c[1]); r.c[2] = fp_neg(a.c[2]); break;
        case OP_MULPOW: r = fq3_mul(a, fq3_pow(b, exponent)); break;
        default: r = a;
        }
        store_el(dst, dfield, i, r);
    }
}
/* dst[i] = lhs[i] OP constant  (MulIntoConst/AddIntoConst/…; FillBuff when op==OP_FILL) */
void orc_pointwise_const(int op, unsigned dfield, u64 *dst, unsigned lfield, const u64 *lhs,
                         unsigned cfield, const u64 *cst, size_t n) {
    fq3 b = load_el(cst, cfield, 0);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        fq3 a = lhs ? load_el(lhs, lfield, i) : fq3_zero(), r;
        switch (op) {
        case OP_MUL: r = fq3_mul(a, b); break;
        case OP_ADD: r = fq3_add(a, b); break;
        case OP_SUB: r = fq3_sub(a, b); break;
        case OP_FILL: r = b; break;
        default: r = a;
        }
        store_el(dst, dfield, i, r);
    }
}

/* acc[i] = sum over columns (Matrix::sum_columns, src/matrix.rs:322-394) */
void orc_sum_columns(const u64 *base, size_t col_stride_words, unsigned ncols, unsigned lanes, size_t n, u64 *acc) {
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n * lanes; i++) {
        u64 s = 0;
        for (unsigned c = 0; c < ncols; c++) s = fp_add(s, base[(size_t)c * col_stride_words + i]);
        acc[i] = s;
    }
}

/* ------------------------------------------------------------------ FRI -- */
/* apply_drp (src/fri.rs:526-567) on bit-reversed-order evaluations over the
 * coset offset*<g_n>: bit_reverse, iNTT, *ff, fold by alpha, NTT(n/ff) on the
 * coset offset^ff, bit_reverse.  evals: n elements of `lanes` words
 * (lanes=3: Fq3 with alpha in Fq3; lanes=1: Fq = Fp).  out: n/ff elements. */
void orc_fri_apply_drp(const u64 *evals, unsigned lanes, unsigned log_n, unsigned log_ff,
                       u64 offset_mont, const u64 *alpha, u64 *out) {
    size_t n = (size_t)1 << log_n, ff = (size_t)1 << log_ff, m = n >> log_ff;
    u64 *co = (u64 *)malloc(sizeof(u64) * n * lanes);
    memcpy(co, evals, sizeof(u64) * n * lanes);
    orc_bit_reverse(co, lanes, log_n);
    orc_ntt_columns(co, 0, 1, lanes, log_n, offset_mont, 1);
    u64 ffm = fp_from_canon((u64)ff);
    fq3 al = load_el(alpha, lanes, 0);
    fq3 apow[64];
    apow[0] = fq3_one();
    for (size_t i = 1; i < ff; i++) apow[i] = fq3_mul(apow[i - 1], al);
    #pragma omp parallel for schedule(static)
    for (size_t k = 0; k < m; k++) {
        fq3 s = fq3_zero();
        for (size_t i = 0; i < ff; i++) {
            fq3 c = fq3_mul_fp(load_el(co, lanes, k * ff + i), ffm);
            s = fq3_add(s, fq3_mul(c, apow[i]));
        }
        store_el(out, lanes, k, s);
    }
    free(co);
    u64 drp_off = fp_pow(offset_mont, ff);
    orc_ntt_columns(out, 0, 1, lanes, log_n - log_ff, drp_off, 0);
    orc_bit_reverse(out, lanes, log_n - log_ff);
}

/* ----------------------------------------------------------------- DEEP -- */
/* horner_evaluate (src/utils.rs:124-131): coeffs in field `cf` (1|3), point in Fq3 */
void orc_horner(const u64 *coeffs, unsigned cf, size_t n, const u64 *point, u64 *out) {
    fq3 x = load_el(point, 3, 0), r = fq3_zero();
    for (size_t i = n; i-- > 0;) r = fq3_add(fq3_mul(r, x), load_el(coeffs, cf, i));
    memcpy(out, &r, 24);
}
/* divide_out_points_into (src/utils.rs:163-175): in place on Fq3 coeffs.
 *   coeff_i <- sum_k c_k * rem_k ;  rem_k <- z_k*rem_k + old coeff_i   (i descending) */
void orc_divide_out_points(u64 *coeffs, size_t n, const u64 *zs, const u64 *cs, unsigned k) {
    fq3 rem[64];
    for (unsigned j = 0; j < k; j++) rem[j] = fq3_zero();
    for (size_t i = n; i-- > 0;) {
        fq3 tmp = load_el(coeffs, 3, i), acc = fq3_zero();
        for (unsigned j = 0; j < k; j++) acc = fq3_add(acc, fq3_mul(load_el(cs, 3, j), rem[j]));
        store_el(coeffs, 3, i, acc);
        for (unsigned j = 0; j < k; j++) rem[j] = fq3_add(fq3_mul(load_el(zs, 3, j), rem[j]), tmp);
    }
}
/* get_ood_evals (src/composer.rs:60-83): one Horner evaluation per (column, point) job, jobs across the threads like the
 * reference's cfg_into_iter over the trace arguments */
void orc_horner_jobs(const u64 *const *coeffs, const unsigned *cf, size_t n, const u64 *points, unsigned njobs, u64 *out) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (unsigned j = 0; j < njobs; j++) orc_horner(coeffs[j], cf[j], n, points + 3 * (size_t)j, out + 3 * (size_t)j);
}
/* into_deep_poly (src/composer.rs:108-158): every column divided by its own points, columns across the threads.
 * cols: ncols columns of n Fq3 coefficients (in place); column c uses counts[c] points from the running lists zs / cs */
void orc_divide_out_points_columns(u64 *base, size_t col_stride_words, unsigned ncols, size_t n, const u64 *zs, const u64 *cs,
                                   const unsigned *counts) {
    size_t *start = (size_t *)malloc(sizeof(size_t) * (ncols + 1));
    start[0] = 0;
    for (unsigned c = 0; c < ncols; c++) start[c + 1] = start[c] + counts[c];
    #pragma omp parallel for schedule(dynamic, 1)
    for (unsigned c = 0; c < ncols; c++)
        orc_divide_out_points(base + (size_t)c * col_stride_words, n, zs + 3 * start[c], cs + 3 * start[c], counts[c]);
    free(start);
}
/* running product / running evaluation columns as the reference builds them row by row
 * (examples/brainfuck/trace.rs:108-279):  x_0 = init, x_(i+1) = x_i * a_i + b_i;
 * out[i] = x_i (exclusive: the value stored BEFORE the row's update) or x_(i+1) (inclusive).
 * a (field fa) NULL => a_i = a_const; b (field fb) NULL => 0; out: n elements of `field`. */
void orc_scan_affine(unsigned field, const u64 *a, unsigned fa, const u64 *a_const, const u64 *b, unsigned fb, size_t n,
                     const u64 *init, int inclusive, u64 *out) {
    fq3 x = load_el(init, 3, 0), ac = a_const ? load_el(a_const, 3, 0) : fq3_one();
    for (size_t i = 0; i < n; i++) {
        if (!inclusive) store_el(out, field, i, x);
        x = fq3_mul(x, a ? load_el(a, fa, i) : ac);
        if (b) x = fq3_add(x, load_el(b, fb, i));
        if (inclusive) store_el(out, field, i, x);
    }
}
/* degree adjustment P(x)*(alpha + beta x) (src/composer.rs:167-185), Fq3 coeffs in place */
void orc_degree_adjust(u64 *coeffs, size_t n, const u64 *alpha, const u64 *beta) {
    fq3 a = load_el(alpha, 3, 0), b = load_el(beta, 3, 0), last = fq3_zero();
    for (size_t i = 0; i < n; i++) {
        fq3 tmp = load_el(coeffs, 3, i);
        store_el(coeffs, 3, i, fq3_add(fq3_mul(tmp, a), fq3_mul(last, b)));
        last = tmp;
    }
}

/* proof of work, serial branch of PublicCoin::grind_proof_of_work (src/random.rs:48-51,129-132) */
static u32 leading_zero_bits(const uint8_t d[32]) {
    u32 z = 0;
    for (int i = 0; i < 32; i++) {
        if (d[i] == 0) { z += 8; continue; }
        z += (u32)__builtin_clz((unsigned)d[i]) - 24;
        break;
    }
    return z;
}
u64 orc_pow_grind(const uint8_t seed[32], unsigned bits) {
    uint8_t msg[40], dig[32];
    memcpy(msg, seed, 32);
    for (u64 nonce = 1;; nonce++) {
        for (int i = 0; i < 8; i++) msg[32 + i] = (uint8_t)(nonce >> (8 * (7 - i)));
        orc_sha256(msg, 40, dig);
        if (leading_zero_bits(dig) >= bits) return nonce;
    }
}

/* --------------------------------------------------------- misc helpers -- */
/* synthetic data: splitmix64, reject >= p, canonical x -> x*2^64 mod p */
void orc_splitmix_fill(u64 *dst, size_t n, u64 seed) {
    u64 s = seed;
    for (size_t i = 0; i < n;) {
        s += 0x9E3779B97F4A7C15ULL;
        u64 z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        if (z < GL_P) dst[i++] = fp_from_canon(z);
    }
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}
