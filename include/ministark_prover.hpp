// ministark_prover.hpp — `default_prove` (src/prover.rs:25-174) in C++ on top of the C ABI: the compiled-language
// counterpart of ministark_b200/prover.py (same transcript order, same calls), for AIRs without extension columns
// (examples/fib; the extension-column hook of the Python driver is not mirrored yet).
//
// Host logic (coin, AIR, programs, wire format) comes from ministark_host.hpp and is CPU-tested.  This file only strings
// the ms_* calls together; it is compile- and link-checked in the build container and exercised on a GPU by
// tests/test_gpu_cpp_prover.py, which compares its proof bytes with the Python driver's.
#pragma once
#include <stdexcept>

#include "ministark_b200.h"
#include "ministark_host.hpp"

namespace mshost {

struct DeviceBuf {   // RAII over ms_alloc_device
    ms_ctx *ctx = nullptr;
    void *p = nullptr;
    DeviceBuf() = default;
    DeviceBuf(ms_ctx *c, size_t bytes) : ctx(c) {
        if (ms_alloc_device(c, bytes, &p) != MS_OK) throw std::runtime_error(std::string("ms_alloc_device: ") + ms_last_error(c));
    }
    DeviceBuf(const DeviceBuf &) = delete;
    DeviceBuf &operator=(const DeviceBuf &) = delete;
    DeviceBuf(DeviceBuf &&o) noexcept : ctx(o.ctx), p(o.p) { o.p = nullptr; }
    DeviceBuf &operator=(DeviceBuf &&o) noexcept {
        if (this != &o) { release(); ctx = o.ctx; p = o.p; o.p = nullptr; }
        return *this;
    }
    ~DeviceBuf() { release(); }
    void release() { if (p) ms_free(ctx, p); p = nullptr; }
    u64 *words() const { return static_cast<u64 *>(p); }
};

inline void ck(ms_ctx *c, int rc, const char *what) {
    if (rc != MS_OK) throw std::runtime_error(std::string(what) + ": " + ms_last_error(c));   // the reference panics
}

// DEEP composition as a symbolic expression over LDE columns [base..., composition...] (ministark_b200/deep.py):
// hints are, in order of first use: composition ood / z^m / composition alpha per column, then per trace argument
// its ood value, its point z*g^offset (shared per offset) and its alpha, then the two degree coefficients.
struct DeepKey { int kind; int64_t index; bool operator<(const DeepKey &o) const { return std::tie(kind, index) < std::tie(o.kind, o.index); } };
enum { DK_COOD, DK_ZM, DK_CALPHA, DK_TOOD, DK_ZPT, DK_TALPHA, DK_DALPHA, DK_DBETA };
inline Expr deep_expression(Graph &g, const std::vector<std::pair<u64, int64_t>> &trace_arguments, u32 num_trace_cols,
                            u32 num_composition_cols, std::vector<DeepKey> &keys) {
    std::map<DeepKey, u64> index;
    auto H = [&](int kind, int64_t i) {
        DeepKey k{kind, i};
        if (!index.count(k)) { index[k] = keys.size(); keys.push_back(k); }
        return Hint(g, index[k]);
    };
    Expr x = X(g), one = Constant(g, 1), total;
    std::map<DeepKey, Expr> inv_cache;
    auto inv_x_minus = [&](int kind, int64_t i) {
        DeepKey k{kind, i};
        if (!inv_cache.count(k)) inv_cache[k] = one / (x - H(kind, i));
        return inv_cache[k];
    };
    bool first = true;
    for (u32 j = 0; j < num_composition_cols; j++) {
        Expr term = (Trace(g, num_trace_cols + j, 0) - H(DK_COOD, j)) * inv_x_minus(DK_ZM, 0) * H(DK_CALPHA, j);
        total = first ? term : total + term;
        first = false;
    }
    for (size_t i = 0; i < trace_arguments.size(); i++) {
        Expr term = (Trace(g, trace_arguments[i].first, 0) - H(DK_TOOD, (int64_t)i)) * inv_x_minus(DK_ZPT, trace_arguments[i].second) * H(DK_TALPHA, (int64_t)i);
        total = first ? term : total + term;
        first = false;
    }
    return total * (H(DK_DALPHA, 0) + x * H(DK_DBETA, 0));
}

class GpuProver {
    ms_ctx *ctx = nullptr;

public:
    explicit GpuProver(int device = 0) {
        if (ms_ctx_create(device, &ctx) != MS_OK) throw std::runtime_error("ms_ctx_create failed (no CUDA device? there is no CPU fallback)");
    }
    ~GpuProver() { if (ctx) ms_ctx_destroy(ctx); }
    GpuProver(const GpuProver &) = delete;
    GpuProver &operator=(const GpuProver &) = delete;

    // base_trace: num_base_columns x n Montgomery words, column-major, HOST memory.  public_inputs: what the claim
    // serialises into the coin seed (examples/fib: the claimed value) and hands to gen_hints.
    Proof prove(const AirConfig &cfg, ProofOptions options, const u64 *base_trace, u64 n, const std::vector<Fq> &public_inputs) {
        if (cfg.num_extension_columns) throw std::runtime_error("extension columns are not supported by the C++ driver yet");
        const int fq = cfg.fq_is_fp ? MS_FIELD_FP : MS_FIELD_FQ3, lanes = cfg.fq_is_fp ? 1 : 3;
        Air air(cfg, n, options);
        const unsigned log_n = air.log_n, beta = options.lde_blowup_factor, log_b = 31 - (unsigned)__builtin_clz(beta);
        const unsigned log_N = log_n + log_b;
        const u64 N = n << log_b, ce = air.ce_blowup_factor, M = n * ce;
        const unsigned log_ce = log_n + (63 - (unsigned)__builtin_clzll(ce));
        const u32 nbase = cfg.num_base_columns;
        const u64 GEN = to_mont(GENERATOR), ONE = to_mont(1);
        // gen_public_coin (examples/fib/main.rs:166-172)
        Bytes seed;
        for (const Fq &v : public_inputs) put_elem(seed, v, lanes);
        put_u64_le(seed, n);
        for (u8 b : options.to_bytes()) seed.push_back(b);
        PublicCoin coin(sha256({seed}), lanes);
        Proof proof;
        proof.options = options;
        proof.trace_len = n;

        // ---- base trace commitment (prover.rs:46-55)
        DeviceBuf d_trace(ctx, (size_t)nbase * n * 8), base_polys(ctx, (size_t)nbase * n * 8), base_lde(ctx, (size_t)nbase * N * 8);
        DeviceBuf base_leaves(ctx, N * 32), base_nodes(ctx, N * 32);
        ck(ctx, ms_copy(ctx, d_trace.p, base_trace, (size_t)nbase * n * 8), "upload");
        ck(ctx, ms_ntt_batch_to(ctx, MS_FIELD_FP, d_trace.p, n, base_polys.p, n, nbase, log_n, MS_NTT_INVERSE, ONE), "interpolate");
        ck(ctx, ms_lde_batch(ctx, MS_FIELD_FP, base_polys.p, n, base_lde.p, N, nbase, log_n, log_b, GEN, 1), "lde");
        proof.base_trace_commitment.resize(32);
        ck(ctx, ms_merkle_commit_sha256(ctx, MS_FIELD_FP, base_lde.p, N, nbase, N, base_leaves.p, base_nodes.p, proof.base_trace_commitment.data()), "commit");
        d_trace.release();
        coin.reseed_with_digest(proof.base_trace_commitment);
        std::vector<Fq> challenges;
        for (u64 i = 0; i < air.num_challenges(); i++) challenges.push_back(coin.draw());
        const std::vector<Fq> hints = cfg.gen_hints ? cfg.gen_hints(n, public_inputs, challenges) : std::vector<Fq>{};

        // ---- constraint evaluation over the ce domain, read in place from the bit-reversed LDE prefix (prover.rs:75-108)
        std::vector<Fq> ccoefs;
        for (u64 i = 0; i < air.num_composition_constraint_coeffs(); i++) ccoefs.push_back(coin.draw());
        const Program prog = air.composition_program(nbase).bind(challenges, hints, ccoefs);
        DeviceBuf comp_evals(ctx, M * lanes * 8);
        ck(ctx, ms_eval_constraints(ctx, &prog.code[0][0], (unsigned)prog.code.size(), &prog.consts[0][0], (unsigned)prog.consts.size(),
                                    base_lde.p, N, nbase, nullptr, N, 0, fq, log_ce, GEN, 1, 0, comp_evals.p), "eval_constraints");

        // ---- composition trace (prover.rs:110-125)
        ck(ctx, ms_ntt_batch(ctx, fq, comp_evals.p, M, 1, log_ce, MS_NTT_INVERSE, GEN), "composition iNTT");
        DeviceBuf comp_split;
        void *comp_polys = comp_evals.p;
        if (ce > 1) {
            comp_split = DeviceBuf(ctx, M * lanes * 8);
            ck(ctx, ms_matrix_from_rows(ctx, fq, comp_evals.p, n, (unsigned)ce, comp_split.p, n), "composition split");
            comp_polys = comp_split.p;
        }
        DeviceBuf comp_lde(ctx, ce * N * lanes * 8), comp_leaves(ctx, N * 32), comp_nodes(ctx, N * 32);
        ck(ctx, ms_lde_batch(ctx, fq, comp_polys, n, comp_lde.p, N, (unsigned)ce, log_n, log_b, GEN, 1), "composition lde");
        proof.composition_trace_commitment.resize(32);
        ck(ctx, ms_merkle_commit_sha256(ctx, fq, comp_lde.p, N, (unsigned)ce, N, comp_leaves.p, comp_nodes.p, proof.composition_trace_commitment.data()),
           "composition commit");
        coin.reseed_with_digest(proof.composition_trace_commitment);

        // ---- out-of-domain evaluations (composer.rs:43-86)
        const Fq z = coin.draw();
        const auto trace_args = air.trace_arguments();
        std::vector<int64_t> offsets;
        for (const auto &ta : trace_args)
            if (std::find(offsets.begin(), offsets.end(), ta.second) == offsets.end()) offsets.push_back(ta.second);
        std::sort(offsets.begin(), offsets.end());
        const u64 g = domain_generator(log_n), g_inv = invm(g);
        std::vector<Fq> z_points;
        std::vector<u64> pts;
        for (int64_t o : offsets) {
            const Fq p = fq_scale(z, powm(o >= 0 ? g : g_inv, (u64)(o >= 0 ? o : -o)));
            z_points.push_back(p);
            for (int l = 0; l < 3; l++) pts.push_back(to_mont(p.c[l]));
        }
        const Fq z_m = fq_pow(z, ce);
        std::vector<u64> base_ood((size_t)nbase * offsets.size() * 3), comp_ood(ce * 3);
        ck(ctx, ms_poly_eval(ctx, MS_FIELD_FP, base_polys.p, n, nbase, n, pts.data(), (unsigned)offsets.size(), base_ood.data()), "ood (trace)");
        const u64 zm_w[3] = {to_mont(z_m.c[0]), to_mont(z_m.c[1]), to_mont(z_m.c[2])};
        ck(ctx, ms_poly_eval(ctx, fq, comp_polys, n, (unsigned)ce, n, zm_w, 1, comp_ood.data()), "ood (composition)");
        auto canon3 = [&](const u64 *w) {
            Fq v(from_mont(w[0]), from_mont(w[1]), from_mont(w[2]));
            if (lanes == 1 && (v.c[1] || v.c[2])) throw std::runtime_error("out-of-domain value left the base field although Fq = Fp");
            return v;
        };
        for (const auto &ta : trace_args) {
            const size_t k = std::find(offsets.begin(), offsets.end(), ta.second) - offsets.begin();
            proof.execution_trace_ood_evals.push_back(canon3(&base_ood[(ta.first * offsets.size() + k) * 3]));
        }
        for (u64 j = 0; j < ce; j++) proof.composition_trace_ood_evals.push_back(canon3(&comp_ood[j * 3]));
        std::vector<Fq> all_oods = proof.execution_trace_ood_evals;
        all_oods.insert(all_oods.end(), proof.composition_trace_ood_evals.begin(), proof.composition_trace_ood_evals.end());
        coin.reseed_with_field_elements(all_oods);

        // ---- DEEP composition evaluated over the LDE domain (composer.rs:89-188 in evaluation form)
        std::vector<Fq> ex_alphas, co_alphas;
        for (size_t i = 0; i < trace_args.size(); i++) ex_alphas.push_back(coin.draw());
        for (u64 j = 0; j < ce; j++) co_alphas.push_back(coin.draw());
        const Fq d_alpha = coin.draw(), d_beta = coin.draw();
        Graph dg;
        std::vector<DeepKey> keys;
        const Expr dexpr = deep_expression(dg, trace_args, nbase, (u32)ce, keys);
        std::vector<Fq> dhints;
        for (const DeepKey &k : keys) {
            switch (k.kind) {
                case DK_COOD: dhints.push_back(proof.composition_trace_ood_evals[k.index]); break;
                case DK_ZM: dhints.push_back(z_m); break;
                case DK_CALPHA: dhints.push_back(co_alphas[k.index]); break;
                case DK_TOOD: dhints.push_back(proof.execution_trace_ood_evals[k.index]); break;
                case DK_ZPT: dhints.push_back(z_points[std::find(offsets.begin(), offsets.end(), k.index) - offsets.begin()]); break;
                case DK_TALPHA: dhints.push_back(ex_alphas[k.index]); break;
                case DK_DALPHA: dhints.push_back(d_alpha); break;
                default: dhints.push_back(d_beta);
            }
        }
        const Program dprog = compile_program(dg, dexpr.id, nbase, 1, (int)log_N).bind({}, dhints, {});
        std::vector<const void *> cols;
        std::vector<int> is_q;
        for (u32 c = 0; c < nbase; c++) { cols.push_back(base_lde.words() + (size_t)c * N); is_q.push_back(0); }
        for (u64 j = 0; j < ce; j++) { cols.push_back(comp_lde.words() + (size_t)j * N * lanes); is_q.push_back(1); }
        DeviceBuf cur(ctx, N * lanes * 8);
        ck(ctx, ms_eval_constraints_ptrs(ctx, &dprog.code[0][0], (unsigned)dprog.code.size(), &dprog.consts[0][0], (unsigned)dprog.consts.size(),
                                         cols.data(), is_q.data(), (unsigned)cols.size(), fq, log_N, GEN, 1, 1, cur.p), "deep composition");

        // ---- FRI (fri.rs:179-249)
        const unsigned ff = options.fri_folding_factor, log_ff = 31 - (unsigned)__builtin_clz(ff);
        struct Layer { DeviceBuf evals, leaves, nodes; Bytes root; u64 nrows; };
        std::vector<Layer> layers;
        unsigned ln = log_N;
        for (unsigned l = 0; l < options.fri_num_layers(N); l++) {
            Layer L;
            L.nrows = (u64)1 << (ln - log_ff);
            L.leaves = DeviceBuf(ctx, L.nrows * 32);
            L.nodes = DeviceBuf(ctx, L.nrows * 32);
            L.root.resize(32);
            ck(ctx, ms_merkle_commit_rows_sha256(ctx, cur.p, ff * lanes, L.nrows, L.leaves.p, L.nodes.p, L.root.data()), "fri layer commit");
            coin.reseed_with_digest(L.root);
            const Fq alpha = coin.draw();
            const u64 aw[3] = {to_mont(alpha.c[0]), to_mont(alpha.c[1]), to_mont(alpha.c[2])};
            DeviceBuf nxt(ctx, L.nrows * lanes * 8);
            ck(ctx, ms_fri_fold(ctx, fq, cur.p, ln, log_ff, ONE, aw, nxt.p), "fri fold");
            L.evals = std::move(cur);
            cur = std::move(nxt);
            layers.push_back(std::move(L));
            ln -= log_ff;
        }
        {   // set_remainder (fri.rs:233-249)
            const u64 rem_size = (u64)1 << ln;
            ck(ctx, ms_bit_reverse(ctx, fq, cur.p, rem_size, 1, ln), "remainder bit reverse");
            ck(ctx, ms_ntt_batch(ctx, fq, cur.p, rem_size, 1, ln, MS_NTT_INVERSE, ONE), "remainder iNTT");
            std::vector<u64> w(rem_size * lanes);
            ck(ctx, ms_copy(ctx, w.data(), cur.p, w.size() * 8), "remainder download");
            const u64 keep = rem_size / beta;
            for (u64 i = 0; i < rem_size; i++) {
                Fq v;
                for (int l = 0; l < lanes; l++) v.c[l] = from_mont(w[i * lanes + l]);
                if (i < keep) proof.fri_proof.remainder_coeffs.push_back(v);
                else if (!v.is_zero()) throw std::runtime_error("FRI remainder is not low degree");
            }
            coin.reseed_with_field_elements(proof.fri_proof.remainder_coeffs);
        }
        // ---- proof of work + queries
        if (options.grinding_factor) {
            ck(ctx, ms_pow_grind_sha256(ctx, coin.seed.data(), options.grinding_factor, &proof.pow_nonce), "pow");
            if (!coin.verify_proof_of_work(options.grinding_factor, proof.pow_nonce)) throw std::runtime_error("bad nonce");
            coin.reseed_with_int(proof.pow_nonce);
        }
        const std::vector<u64> positions = coin.draw_queries(options.num_queries, N);
        auto view_of = [&](const DeviceBuf &leaves, const DeviceBuf &nodes, u64 nleaves, const std::vector<u64> &idx) {
            const unsigned height = 63 - (unsigned)__builtin_clzll(nleaves);
            std::vector<u8> init(idx.size() * 32), sib(idx.size() * 32), path(idx.size() * (height ? height : 1) * 32);
            unsigned counts[3];
            ck(ctx, ms_merkle_prove_sha256(ctx, leaves.p, nodes.p, nleaves, idx.data(), (unsigned)idx.size(), init.data(), sib.data(), path.data(), counts),
               "merkle prove");
            MerkleView v;
            auto take = [](const std::vector<u8> &b, unsigned k) { std::vector<Bytes> o; for (unsigned i = 0; i < k; i++) o.emplace_back(b.begin() + 32 * i, b.begin() + 32 * i + 32); return o; };
            v.initial_leaves = take(init, counts[0]);
            v.sibling_leaves = take(sib, counts[1]);
            v.nodes = take(path, counts[2]);
            v.height = height;
            return v;
        };
        auto canon_vec = [&](const std::vector<u64> &w, int l) {
            std::vector<Fq> o;
            for (size_t i = 0; i < w.size(); i += l) {
                Fq v;
                for (int k = 0; k < l; k++) v.c[k] = from_mont(w[i + k]);
                o.push_back(v);
            }
            return o;
        };
        std::vector<u64> folded = positions;
        for (Layer &L : layers) {
            std::set<u64> s;
            for (u64 p : folded) s.insert(p / ff);
            folded.assign(s.begin(), s.end());
            std::vector<u64> rows(folded.size() * ff * lanes);
            ck(ctx, ms_gather_rows_rowmajor(ctx, L.evals.p, ff * lanes, L.nrows, folded.data(), (unsigned)folded.size(), rows.data()), "fri rows");
            LayerProof lp;
            lp.flattenend_rows = canon_vec(rows, lanes);
            lp.merkle_proof = view_of(L.leaves, L.nodes, L.nrows, folded);
            lp.commitment = L.root;
            proof.fri_proof.layers.push_back(std::move(lp));
        }
        std::vector<u64> brow(positions.size() * nbase), crow(positions.size() * ce * lanes);
        ck(ctx, ms_gather_rows(ctx, MS_FIELD_FP, base_lde.p, N, nbase, N, positions.data(), (unsigned)positions.size(), brow.data()), "base rows");
        ck(ctx, ms_gather_rows(ctx, fq, comp_lde.p, N, (unsigned)ce, N, positions.data(), (unsigned)positions.size(), crow.data()), "composition rows");
        proof.trace_queries.base_trace_values = canon_vec(brow, 1);
        proof.trace_queries.composition_trace_values = canon_vec(crow, lanes);
        proof.trace_queries.base_trace_proof = view_of(base_leaves, base_nodes, N, positions);
        proof.trace_queries.composition_trace_proof = view_of(comp_leaves, comp_nodes, N, positions);
        return proof;
    }
};

}  // namespace mshost
