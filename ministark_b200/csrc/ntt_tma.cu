// ntt_tma.cu — persistent, warp-specialised TMA pipeline for the 256 x 16 tile passes of the NTT engine.
//
// The large transforms (2^16, 2^24 points per coset: the LDE of a 2^24-row trace is 32 columns x 8 cosets of these)
// are made of passes whose CTA tile is 256 transform rows x 16 lanes = 4096 words = 32 KiB (ntt.cu).  ntt_pass_kernel
// runs one tile per CTA: load -> butterflies -> exchange -> store, and relies on three resident CTAs to overlap those
// phases (ncu, round 1: issue slots 67 % used, ALU pipe idle whenever the three CTAs sit in a memory phase together).
// Here ONE CTA per SM stays resident for the whole pass and is split by role:
//
//   warp 0, lane 0   service thread: issues the TMA loads (cp.async.bulk.tensor, 4-D tensor maps, 128-byte swizzle)
//                    of the next tiles into a ring of 32 KiB shared-memory stages, each guarded by an mbarrier, and
//                    drains finished stages with TMA stores (bulk async-groups).  No thread computes an address of
//                    global memory or touches LDG/STG: the ALU pipe is left to the butterflies.
//   warps 1..8G      consumers: G groups of 8 warps, each group works on its own tile (256 threads x 16 elements), in
//                    place in the stage.  Strided tiles: a half warp spans the 16 lanes of one tile row, so both
//                    radix-16 steps touch whole 128-byte rows — conflict free, no padding, no second copy — and the
//                    exchange between the steps costs one named barrier of the group.  Contiguous tiles: the 16
//                    threads that exchange sit in one half warp (__syncwarp only) and the 128-byte swizzle makes both
//                    the stride-16 step and the whole-row 128-bit step conflict free.  One group's exchange and
//                    mbarrier waits are covered by the other groups' arithmetic.
//
//   pass kinds      strided   rows S words apart (S = 2^16 or 2^8), lanes contiguous: box 16 x 256 of the
//                             (S, 256, blocks*cosets, columns) tensor.  Inter-pass twiddles w_{RS}^(i_R * lower) come
//                             from the plan's full table as a 32 KiB tile loaded once per tile index and shared by
//                             every column and coset; the coset pre-scale q^j (first pass of an LDE) as a second tile
//                             per (tile index, coset), double buffered.
//                   contiguous (last pass of a bit-reversed LDE): 16 sub-problems of 256 consecutive words; the tile
//                             is the same 256 x 128 B box of the (16, rows, columns) view, the second step reads
//                             and writes whole 128-byte rows with 128-bit shared accesses.
//
// Replaces gpu/src/plan.rs:427-450 + fft_shaders.h.metal:61-101 for these shapes; results are bit-identical to
// ntt_pass_kernel (same networks, same tables, same lazy arithmetic) — tests/test_gpu_ntt.py compares both paths.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "dft.cuh"
#include "ntt.cuh"

namespace msntt {

using namespace gl;

namespace {

constexpr int kTileWords = 4096;
constexpr u32 kTileBytes = 32768;
constexpr int kMaxStages = 8;

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void tma_load4(u32 dst, const CUtensorMap *tm, u32 bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap *tm, u32 src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<unsigned long long>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(tm)) : "memory");
}

// word offset of (row, element e in 0..15) inside a 256 x 128 B tile written by TMA with the 128-byte swizzle:
// the 16-byte chunk index is XORed with row % 8.
__host__ __device__ constexpr u32 swz(u32 row, u32 e) { return row * 16 + ((((e >> 1) ^ (row & 7))) << 1) + (e & 1); }

constexpr int kMaxScatterCos = 16;
// LDE scatter (multi-GPU fused exchange): one tensor map per destination coset block (possibly PEER memory mapped with
// CUDA IPC: the TMA store goes out over NVLink) and per optional second copy
struct ScatterMaps {
    CUtensorMap out[kMaxScatterCos];
    CUtensorMap dup[kMaxScatterCos];
    u32 dup_mask;        // bit q: block q has a second destination
    u32 log_chunks;      // log2(4096-word chunks per coset block)
};

struct TmaArgs {
    u32 ncols;
    u32 log_ncos;     // cosets per column (LDE), 0 otherwise
    u32 units;        // strided: tiles * cosets (coset fastest); contiguous: 4096-word chunks per column
    u32 log_ext0;     // strided: log2(S / 16)
    u32 nblk;         // strided: N / (256 * S)
    u32 in_has_cos;   // strided: the input tensor has a coset dimension (every pass but the first of an LDE)
    u32 nstages;
    const u64 *t4096; // omega_4096^e of the transform's direction
};

// inter-step twiddles of the 256-point sub-NTT: x[brev(KAP)] *= omega_256^(KAP * a); row a of t16 holds the 16
// factors, 16-byte chunks XOR-swizzled by a % 8 so that the 8 lanes of a quarter warp read 8 different chunks
template <bool INV>
__device__ __forceinline__ void inner_twiddles(u64 (&x)[16], const u64 *__restrict__ t16, const u32 a) {
    const ulonglong2 *trow = reinterpret_cast<const ulonglong2 *>(t16 + a * 16);
    static_for<0, 8>([&](auto J) {
        const ulonglong2 w = trow[(u32)J ^ (a & 7)];
        constexpr int k0 = 2 * decltype(J)::value, k1 = k0 + 1;
        if constexpr (k0 > 0) x[brev_c(k0, 4)] = mul(x[brev_c(k0, 4)], w.x);
        else x[0] = canon(x[0]);
        x[brev_c(k1, 4)] = mul(x[brev_c(k1, 4)], w.y);
    });
}

__device__ __forceinline__ void group_sync(const u32 g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

// One thread's share of a strided-pass tile, in place.  tg = thread index inside the consumer group (256 threads): the
// 16 lanes of a half warp are the 16 lanes of ONE tile row, so every access below reads or writes whole 128-byte rows
// (2 rows per warp access = 2 wavefronts, the minimum) whatever the row: both radix-16 steps are conflict free.  (Lanes
// running over the rows instead would make the exchange warp-local, but rows 16 b + K share row % 8 — the only input of
// the 128-byte swizzle — and step 2 would serialise 16-fold.)  The exchange between the steps spans the group: one
// named barrier.
template <bool INV, bool BITREV, bool HAS_PRE>
__device__ __forceinline__ void tile_strided(u64 *__restrict__ tile, const u64 *__restrict__ tw, const u64 *__restrict__ pre,
                                             const u64 *__restrict__ t16, const u32 tg, const u32 g) {
    const u32 c = tg & 15;
    const u32 a = tg >> 4;
    u64 x[16];
    {   // step 1: rows a + 16 K (high digit), outputs kappa1 -> rows a + 16 * (brev) kappa1
        const u32 o1 = swz(a, c);   // (a + 16 K) % 8 == a % 8: one swizzle term for all K
        static_for<0, 16>([&](auto K) { x[K] = tile[o1 + 256 * (u32)K]; });
        if constexpr (HAS_PRE) static_for<0, 16>([&](auto K) { x[K] = mul(x[K], pre[o1 + 256 * (u32)K]); });
        dft_regs<4, INV>(x);
        inner_twiddles<INV>(x, t16, a);
        static_for<0, 16>([&](auto Q) { tile[o1 + 256 * (u32)Q] = BITREV ? x[Q] : x[brev_c(decltype(Q)::value, 4)]; });
    }
    group_sync(g);
    {   // step 2: rows 16 b + K (low digit); i_R = kappa1 + 16 kappa2
        const u32 b = a;
        const u32 o2 = 256 * b + c;   // row 16 b + K has swizzle term K % 8: offset = (o2 ^ ((K & 7) << 1)) + 16 K
        static_for<0, 16>([&](auto K) { x[K] = tile[(o2 ^ (((u32)K & 7) << 1)) + 16 * (u32)K]; });
        dft_regs<4, INV>(x);
        const u32 k1 = BITREV ? (__brev(b) >> 28) : b;
        const u32 ot = swz(k1, c);
        static_for<0, 16>([&](auto KAP) {
            constexpr int q = brev_c(decltype(KAP)::value, 4);
            x[q] = mul(x[q], tw[ot + 256 * (u32)KAP]);
        });
        if constexpr (BITREV) {
            static_for<0, 16>([&](auto Q) { tile[(o2 ^ (((u32)Q & 7) << 1)) + 16 * (u32)Q] = x[Q]; });
        } else {
            group_sync(g);   // rows b + 16 kappa2 are other threads' inputs: everyone has read before anyone writes
            const u32 on = swz(b, c);
            static_for<0, 16>([&](auto KAP) { tile[on + 256 * (u32)KAP] = x[brev_c(decltype(KAP)::value, 4)]; });
        }
    }
}

// One warp's share (sub-problems c, c+1: 256 consecutive words each) of a contiguous-pass tile (bit-reversed digit).
template <bool INV>
__device__ __forceinline__ void tile_contig(u64 *__restrict__ tile, const u64 *__restrict__ t16, const u32 lane, const u32 wp) {
    const u32 c = 2 * wp + (lane >> 4);
    const u32 a = lane & 15;
    u64 x[16];
    {   // step 1: words a + 16 K of sub-problem c: TMA row 16 c + K, element a
        const u32 o1 = 256 * c + a;
        static_for<0, 16>([&](auto K) { x[K] = tile[(o1 ^ (((u32)K & 7) << 1)) + 16 * (u32)K]; });
        dft_regs<4, INV>(x);
        inner_twiddles<INV>(x, t16, a);
        static_for<0, 16>([&](auto Q) { tile[(o1 ^ (((u32)Q & 7) << 1)) + 16 * (u32)Q] = x[Q]; });
    }
    __syncwarp();
    {   // step 2: words 16 b .. 16 b + 15 = one 128-byte row, chunk j stored at j ^ (b % 8)
        const u32 b = a;
        ulonglong2 *row = reinterpret_cast<ulonglong2 *>(tile + (16 * c + b) * 16);
        static_for<0, 8>([&](auto J) {
            const ulonglong2 v = row[(u32)J ^ (b & 7)];
            x[2 * decltype(J)::value] = v.x;
            x[2 * decltype(J)::value + 1] = v.y;
        });
        dft_regs<4, INV>(x);
        static_for<0, 16>([&](auto Q) { x[Q] = canon(x[Q]); });
        static_for<0, 8>([&](auto J) {
            row[(u32)J ^ (b & 7)] = make_ulonglong2(x[2 * decltype(J)::value], x[2 * decltype(J)::value + 1]);
        });
    }
}

struct Coord {
    int c0, c1, c2, c3;
};
template <int TYPE>
__device__ __forceinline__ void coords(const TmaArgs &A, u32 u, u32 col, Coord &in, Coord &out, Coord &twc, Coord &prec) {
    if constexpr (TYPE == 0) {
        const u32 cos = u & ((1u << A.log_ncos) - 1), t = u >> A.log_ncos;
        const u32 idx0 = t & ((1u << A.log_ext0) - 1), idx1 = t >> A.log_ext0;
        in = Coord{(int)(16 * idx0), 0, (int)((A.in_has_cos ? cos * A.nblk : 0) + idx1), (int)col};
        out = Coord{(int)(16 * idx0), 0, (int)(cos * A.nblk + idx1), (int)col};
        twc = Coord{(int)(16 * idx0), 0, 0, 0};
        prec = Coord{(int)(16 * idx0), 0, (int)(cos * A.nblk + idx1), 0};
    } else {
        in = out = Coord{0, (int)(256 * u), (int)col, 0};
        twc = prec = Coord{0, 0, 0, 0};
    }
}

// TYPE 0: strided pass, 1: contiguous (last, bit-reversed) pass.  G consumer groups of 8 warps.
// scat != nullptr (contiguous pass only): every 4096-word chunk is stored through the tensor map of its coset block.
template <int TYPE, bool INV, bool BITREV, bool HAS_PRE, int G>
__device__ __forceinline__ void ntt_tma_body(const CUtensorMap &tm_in, const CUtensorMap &tm_out, const CUtensorMap &tm_tw,
                                             const CUtensorMap &tm_pre, const TmaArgs &A, const ScatterMaps *scat) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr bool HAS_TW = TYPE == 0;
    constexpr u32 NW = 8 * G;
    const u32 NS = A.nstages;
    u64 *stages = reinterpret_cast<u64 *>(smem_raw);
    u64 *tw_s = stages + (size_t)NS * kTileWords;
    u64 *pre_s = tw_s + (HAS_TW ? kTileWords : 0);
    u64 *t16 = pre_s + (HAS_PRE ? 2 * kTileWords : 0);
    u64 *bars = t16 + 256;
    // barriers: full[NS], computed[NS], tw_full, tw_empty, pre_full[2], pre_empty[2]
    const u32 bar0 = smem_u32(bars);
    auto full_bar = [&](u32 s) { return bar0 + 8 * s; };
    auto comp_bar = [&](u32 s) { return bar0 + 8 * (kMaxStages + s); };
    const u32 tw_full = bar0 + 8 * (2 * kMaxStages), tw_empty = tw_full + 8;
    auto pre_full = [&](u32 s) { return bar0 + 8 * (2 * kMaxStages + 2 + s); };
    auto pre_empty = [&](u32 s) { return bar0 + 8 * (2 * kMaxStages + 4 + s); };

    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        if (smem_u32(smem_raw) & 1023u) __trap();   // the 128-byte swizzle needs 1024-byte aligned tiles
        for (u32 s = 0; s < NS; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(comp_bar(s), 8);
        }
        mbar_init(tw_full, 1);
        mbar_init(tw_empty, NW);
        for (u32 s = 0; s < 2; s++) {
            mbar_init(pre_full(s), 1);
            mbar_init(pre_empty(s), NW);
        }
        fence_mbar_init();
        prefetch_tmap(&tm_in);
        prefetch_tmap(&tm_out);
        if (HAS_TW) prefetch_tmap(&tm_tw);
        if (HAS_PRE) prefetch_tmap(&tm_pre);
    }
    if (tid >= 32 && tid < 32 + 256) {
        // t16[a][KAP] = omega_256^(KAP * a) = omega_4096^(16 KAP a), chunks swizzled by a % 8
        const u32 i = tid - 32, a = i >> 4, k = i & 15;
        t16[a * 16 + ((((k >> 1) ^ (a & 7))) << 1) + (k & 1)] = A.t4096[((k * a) & 255) << 4];
    }
    __syncthreads();

    const u32 nct = gridDim.x, bid = blockIdx.x;
    const u32 u0 = (u32)(((u64)A.units * bid) / nct), u1 = (u32)(((u64)A.units * (bid + 1)) / nct);
    const u32 ncols = A.ncols;
    const u32 total = (u1 - u0) * ncols;

    if (warp == 0) {
        if (lane != 0) return;
        // ---- service thread: loads ahead, stores behind
        u32 li = 0, lu = u0, lcol = 0, ls = 0, twj = 0;
        int loaded_t = -1;
        bool need_aux = true;
        u32 si = 0, su = u0, scol = 0, ss = 0, sk = 0, sdone = 0;
        while (si < total) {
            bool progress = false;
            if (li < total && li < sdone + NS) {
                bool ok = true;
                Coord ci, co, ct, cp;
                coords<TYPE>(A, lu, lcol, ci, co, ct, cp);
                if (need_aux) {
                    if constexpr (HAS_TW) {
                        const int t = (int)(lu >> A.log_ncos);
                        if (t != loaded_t) {
                            if (mbar_test(tw_empty, (twj & 1) ^ 1)) {
                                mbar_arrive_expect_tx(tw_full, kTileBytes);
                                tma_load4(smem_u32(tw_s), &tm_tw, tw_full, ct.c0, ct.c1, ct.c2, ct.c3);
                                twj++;
                                loaded_t = t;
                                progress = true;
                            } else {
                                ok = false;
                            }
                        }
                    }
                    if constexpr (HAS_PRE) {
                        if (ok) {
                            const u32 uc = lu - u0, slot = uc & 1;
                            if (mbar_test(pre_empty(slot), ((uc >> 1) & 1) ^ 1)) {
                                mbar_arrive_expect_tx(pre_full(slot), kTileBytes);
                                tma_load4(smem_u32(pre_s + slot * kTileWords), &tm_pre, pre_full(slot), cp.c0, cp.c1, cp.c2, cp.c3);
                                progress = true;
                            } else {
                                ok = false;
                            }
                        }
                    }
                    if (ok) need_aux = false;
                }
                if (ok) {
                    mbar_arrive_expect_tx(full_bar(ls), kTileBytes);
                    tma_load4(smem_u32(stages + (size_t)ls * kTileWords), &tm_in, full_bar(ls), ci.c0, ci.c1, ci.c2, ci.c3);
                    li++;
                    if (++ls == NS) ls = 0;
                    if (++lcol == ncols) {
                        lcol = 0;
                        lu++;
                        need_aux = true;
                    }
                    progress = true;
                }
            }
            if (si < li && mbar_test(comp_bar(ss), sk & 1)) {
                Coord ci, co, ct, cp;
                coords<TYPE>(A, su, scol, ci, co, ct, cp);
                if (scat) {
                    const u32 cos = su >> scat->log_chunks, within = su & ((1u << scat->log_chunks) - 1);
                    tma_store4(&scat->out[cos], smem_u32(stages + (size_t)ss * kTileWords), 0, (int)(256 * within), (int)scol, 0);
                    if ((scat->dup_mask >> cos) & 1)
                        tma_store4(&scat->dup[cos], smem_u32(stages + (size_t)ss * kTileWords), 0, (int)(256 * within), (int)scol, 0);
                } else {
                    tma_store4(&tm_out, smem_u32(stages + (size_t)ss * kTileWords), co.c0, co.c1, co.c2, co.c3);
                }
                tma_commit();
                tma_wait_read<1>();   // every store but the one just issued has finished reading its stage
                sdone = si;
                si++;
                if (++ss == NS) {
                    ss = 0;
                    sk++;
                }
                if (++scol == ncols) {
                    scol = 0;
                    su++;
                }
                progress = true;
            }
            if (!progress) __nanosleep(40);
        }
        tma_wait_all();
        return;
    }

    // ---- consumers
    const u32 cw = warp - 1, g = cw >> 3, wp = cw & 7;
    u32 i = 0, s = 0, k = 0, gi = 0;      // item counter, its stage and wrap count, item index modulo G
    u32 twj = 0;
    int cur_t = -1;
    for (u32 u = u0, uc = 0; u < u1; u++, uc++) {
        const u32 slot = uc & 1;
        if constexpr (HAS_TW) {
            const int t = (int)(u >> A.log_ncos);
            if (t != cur_t) {
                mbar_wait(tw_full, twj & 1);
                twj++;
                cur_t = t;
            }
        }
        if constexpr (HAS_PRE) mbar_wait(pre_full(slot), (uc >> 1) & 1);
        for (u32 col = 0; col < ncols; col++) {
            if (gi == g) {
                mbar_wait(full_bar(s), k & 1);
                u64 *tile = stages + (size_t)s * kTileWords;
                if constexpr (TYPE == 0)
                    tile_strided<INV, BITREV, HAS_PRE>(tile, tw_s, pre_s + slot * kTileWords, t16, wp * 32 + lane, g);
                else
                    tile_contig<INV>(tile, t16, lane, wp);
                fence_proxy_async();   // generic-proxy writes of the tile before the async-proxy (TMA store) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(comp_bar(s));
            }
            i++;
            if (++s == NS) {
                s = 0;
                k++;
            }
            if (++gi == G) gi = 0;
        }
        __syncwarp();
        if (lane == 0) {
            if constexpr (HAS_PRE) mbar_arrive(pre_empty(slot));
            if constexpr (HAS_TW) {
                if (u + 1 == u1 || (int)((u + 1) >> A.log_ncos) != cur_t) mbar_arrive(tw_empty);
            }
        }
    }
    (void)i;
}

template <int TYPE, bool INV, bool BITREV, bool HAS_PRE, int G>
__global__ void __launch_bounds__(32 + 256 * G, 1)
ntt_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_tw, const __grid_constant__ CUtensorMap tm_pre, const TmaArgs A) {
    ntt_tma_body<TYPE, INV, BITREV, HAS_PRE, G>(tm_in, tm_out, tm_tw, tm_pre, A, nullptr);
}

// the last pass of a multi-GPU LDE: contiguous tiles in, one destination per coset block out
template <int G>
__global__ void __launch_bounds__(32 + 256 * G, 1)
ntt_tma_scatter_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ ScatterMaps scat, const TmaArgs A) {
    ntt_tma_body<1, false, true, false, G>(tm_in, tm_in, tm_in, tm_in, A, &scat);
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

// 4-D u64 tensor, box 16 x 256 x 1 x 1, 128-byte swizzle.  dims[0] is the contiguous dimension; strides in words.
bool make_map(CUtensorMap *m, const void *base, const u64 dims[4], const u64 strides_words[3]) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    cuuint64_t gd[4], gs[3];
    for (int i = 0; i < 4; i++) {
        gd[i] = dims[i];
        if (dims[i] == 0 || dims[i] > 0xFFFFFFFFull) return false;
    }
    for (int i = 0; i < 3; i++) {
        gs[i] = strides_words[i] * 8;
        if (gs[i] == 0 || (gs[i] & 15) || gs[i] >= (1ull << 40)) return false;
    }
    const cuuint32_t box[4] = {16, 256, 1, 1}, es[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(base), gd, gs, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int g_groups = 2;     // consumer groups per CTA (2 or 3)
int g_enabled = 1;
int g_max_stages = kMaxStages;

template <int TYPE, bool INV, bool BITREV, bool HAS_PRE, int G>
bool launch_inst(const CUtensorMap &mi, const CUtensorMap &mo, const CUtensorMap &mt, const CUtensorMap &mp, TmaArgs A,
                 cudaStream_t stream) {
    auto kern = ntt_tma_kernel<TYPE, INV, BITREV, HAS_PRE, G>;
    int dev = 0, max_smem = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t fixed = (size_t)(TYPE == 0 ? kTileBytes : 0) + (HAS_PRE ? 2 * kTileBytes : 0) + 2048 + 8 * (2 * kMaxStages + 6);
    if ((size_t)max_smem < fixed + 3 * (size_t)kTileBytes) return false;
    int ns = (int)(((size_t)max_smem - fixed) / kTileBytes);
    if (ns > g_max_stages) ns = g_max_stages;
    if (ns > kMaxStages) ns = kMaxStages;
    if (ns < 3) return false;
    A.nstages = (u32)ns;
    const size_t smem = fixed + (size_t)ns * kTileBytes;
    static bool attr_set[64] = {false};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    unsigned grid = (unsigned)sms;
    if (A.units < grid) grid = A.units;
    kern<<<grid, 32 + 256 * G, smem, stream>>>(mi, mo, mt, mp, A);
    return true;
}

template <int G>
bool launch_scatter(const CUtensorMap &mi, const ScatterMaps &sm, TmaArgs A, cudaStream_t stream) {
    auto kern = ntt_tma_scatter_kernel<G>;
    int dev = 0, max_smem = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t fixed = 2048 + 8 * (2 * kMaxStages + 6);
    int ns = (int)(((size_t)max_smem - fixed) / kTileBytes);
    if (ns > g_max_stages) ns = g_max_stages;
    if (ns > kMaxStages) ns = kMaxStages;
    if (ns < 3) return false;
    A.nstages = (u32)ns;
    static bool attr_set[64] = {false};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    unsigned grid = (unsigned)sms;
    if (A.units < grid) grid = A.units;
    kern<<<grid, 32 + 256 * G, fixed + (size_t)ns * kTileBytes, stream>>>(mi, sm, A);
    return true;
}

template <int TYPE, bool INV, bool BITREV, bool HAS_PRE>
bool launch_g(const CUtensorMap &mi, const CUtensorMap &mo, const CUtensorMap &mt, const CUtensorMap &mp, const TmaArgs &A,
              cudaStream_t stream) {
    if (g_groups == 3) return launch_inst<TYPE, INV, BITREV, HAS_PRE, 3>(mi, mo, mt, mp, A, stream);
    return launch_inst<TYPE, INV, BITREV, HAS_PRE, 2>(mi, mo, mt, mp, A, stream);
}

}  // namespace

void tma_configure(int enabled, int groups, int max_stages) {
    if (enabled >= 0) g_enabled = enabled;
    if (groups == 2 || groups == 3) g_groups = groups;
    if (max_stages >= 3 && max_stages <= kMaxStages) g_max_stages = max_stages;
}

// Runs pass `p` through the TMA pipeline if its shape is one the pipeline covers; false = not handled (nothing was
// launched, the caller uses ntt_pass_kernel).
bool launch_pass_tma(const PassParams &p, const Tables &t, bool inverse, const u64 *in, u64 *out, unsigned ntiles,
                     unsigned ncols, cudaStream_t stream) {
    if (!g_enabled) return false;
    static const bool env_off = [] { const char *e = getenv("MS_NTT_TMA"); return e && e[0] == '0'; }();
    if (env_off) return false;
    if (p.log_r != 8 || p.log_w != 4 || p.estride != 1 || p.lanes != 1 || p.has_post) return false;
    if (p.out_cos_ptr && !p.host_cos_ptr) return false;
    static const bool scatter_off = [] { const char *e = getenv("MS_NTT_TMA_SCATTER"); return e && e[0] == '0'; }();
    if (p.out_cos_ptr && scatter_off) return false;
    const u64 N = p.n_mask + 1;
    const u32 ncos = p.ncos;
    if ((u64)ntiles * ncos * ncols < 1024) return false;   // too little work for 148 persistent CTAs
    CUtensorMap mi, mo, mt, mp;
    memset(&mt, 0, sizeof mt);
    memset(&mp, 0, sizeof mp);
    TmaArgs A;
    memset(&A, 0, sizeof A);
    A.ncols = ncols;
    A.log_ncos = p.log_ncos;
    A.t4096 = t.t4096;
    const bool strided = !p.in_r_fast && !p.out_r_fast;
    const bool contig = p.in_r_fast && p.out_r_fast;
    if (strided) {
        if (!p.has_outer || !p.outer_tab || p.in_cs != 1 || p.out_cs != 1 || p.in_rs != p.out_rs || p.ndims != 2) return false;
        if (p.has_pre && !p.pre_tab) return false;
        const u64 S = p.in_rs;
        if (S < 16 || p.outer_S != S) return false;
        const u64 nblk = N / (256 * S);
        if (p.dims[0].ext != S / 16 || p.dims[1].ext != nblk) return false;
        const bool in_cos = p.in_cos_stride != 0;
        if (in_cos && p.in_cos_stride != N) return false;
        if (p.out_cos_stride != N && ncos > 1) return false;
        A.units = ntiles * ncos;
        A.log_ext0 = p.dims[0].log_ext;
        A.nblk = (u32)nblk;
        A.in_has_cos = in_cos ? 1 : 0;
        const u64 din[4] = {S, 256, in_cos ? nblk * ncos : nblk, ncols};
        const u64 sin_[3] = {S, 256 * S, p.in_col_stride ? p.in_col_stride : N * ncos};
        const u64 dout[4] = {S, 256, nblk * ncos, ncols};
        const u64 sout[3] = {S, 256 * S, p.out_col_stride ? p.out_col_stride : N * ncos};
        const u64 dtw[4] = {S, 256, 1, 1};
        const u64 stw[3] = {S, 256 * S, 256 * S};
        if (!make_map(&mi, in, din, sin_) || !make_map(&mo, out, dout, sout) || !make_map(&mt, p.outer_tab, dtw, stw)) return false;
        if (p.has_pre) {
            if (p.pre_cos_stride != N) return false;
            const u64 dpre[4] = {S, 256, nblk * ncos, 1};
            const u64 spre[3] = {S, 256 * S, N * ncos};
            if (!make_map(&mp, p.pre_tab, dpre, spre)) return false;
        }
        const bool br = p.bitrev_digit != 0;
        if (inverse) {
            if (p.has_pre) return false;
            return br ? launch_g<0, true, true, false>(mi, mo, mt, mp, A, stream)
                      : launch_g<0, true, false, false>(mi, mo, mt, mp, A, stream);
        }
        if (p.has_pre)
            return br ? launch_g<0, false, true, true>(mi, mo, mt, mp, A, stream)
                      : launch_g<0, false, false, true>(mi, mo, mt, mp, A, stream);
        return br ? launch_g<0, false, true, false>(mi, mo, mt, mp, A, stream)
                  : launch_g<0, false, false, false>(mi, mo, mt, mp, A, stream);
    }
    if (contig) {
        if (!p.bitrev_digit || p.has_outer || p.has_pre || inverse || p.in_rs != 1 || p.out_rs != 1 || in != out) return false;
        if (!p.out_cos_ptr && p.in_col_stride != p.out_col_stride) return false;
        if (ncos > 1 && (p.in_cos_stride != N || (!p.out_cos_ptr && p.out_cos_stride != N))) return false;
        const u64 words = N * ncos;   // per column
        A.units = (u32)(words / 4096);
        A.log_ncos = 0;
        if (p.out_cos_ptr) {
            // scatter: chunks are read from the work buffer (in) and stored through per-block tensor maps
            if (ncos > (u32)kMaxScatterCos || N < 4096 || !p.out_col_stride) return false;
            {
                // TMA stores of whole 32 KiB tiles into PEER memory hold their ring stage until the link has taken the
                // tile; measured on the config-3 step: with half of the blocks remote (2 GPUs) the TMA
                // scatter wins (144.8 vs 147.2 ms), with 7/8 remote (8 GPUs) the one-tile-per-CTA kernel, whose many
                // warps keep more stores in flight, does (157.6 vs 160.2 ms).  So: TMA scatter while at most half of the
                // destination blocks live on another device.
                int dev = 0;
                cudaGetDevice(&dev);
                u32 remote = 0;
                for (u32 q = 0; q < ncos; q++) {
                    cudaPointerAttributes at;
                    if (cudaPointerGetAttributes(&at, p.host_cos_ptr[q]) != cudaSuccess) {
                        cudaGetLastError();
                        remote++;
                    } else if (at.device != dev) {
                        remote++;
                    }
                }
                static const bool force = [] { const char *e = getenv("MS_NTT_TMA_SCATTER"); return e && e[0] == '1'; }();
                if (2 * remote > ncos && !force) return false;
            }
            ScatterMaps sm;                  // 4 KiB of tensor maps, copied into the launch as a __grid_constant__
            memset(&sm, 0, sizeof sm);
            u32 lc = 0;
            while ((4096ull << lc) < N) lc++;
            sm.log_chunks = lc;
            const u64 dblk[4] = {16, N / 16, ncols, 1};
            const u64 sblk[3] = {16, p.out_col_stride, p.out_col_stride};
            const u64 sdup[3] = {16, p.dup_col_stride ? p.dup_col_stride : N, p.dup_col_stride ? p.dup_col_stride : N};
            for (u32 q = 0; q < ncos; q++) {
                if (!p.host_cos_ptr[q] || !make_map(&sm.out[q], p.host_cos_ptr[q], dblk, sblk)) return false;
                if (p.host_dup_ptr && p.host_dup_ptr[q]) {
                    if (!make_map(&sm.dup[q], p.host_dup_ptr[q], dblk, sdup)) return false;
                    sm.dup_mask |= 1u << q;
                }
            }
            const u64 din[4] = {16, words / 16, ncols, 1};
            const u64 sin_[3] = {16, p.in_col_stride ? p.in_col_stride : words, p.in_col_stride ? p.in_col_stride : words};
            if (!make_map(&mi, in, din, sin_)) return false;
            return g_groups == 3 ? launch_scatter<3>(mi, sm, A, stream) : launch_scatter<2>(mi, sm, A, stream);
        }
        const u64 d[4] = {16, words / 16, ncols, 1};
        const u64 s[3] = {16, p.in_col_stride ? p.in_col_stride : words, p.in_col_stride ? p.in_col_stride : words};
        if (!make_map(&mi, in, d, s) || !make_map(&mo, out, d, s)) return false;
        return launch_g<1, false, true, false>(mi, mo, mi, mi, A, stream);
    }
    return false;
}

}  // namespace msntt
