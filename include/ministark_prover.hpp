// ministark_prover.hpp — `default_prove` (src/prover.rs:25-174) in C++ on top of the C ABI: the compiled-language
// counterpart of ministark_b200/prover.py (same transcript order, same calls), including the extension-trace phase
// (prover.rs:56-72) through a device-side column builder; bf::device_extension below builds the nine brainfuck
// extension columns from the resident base trace with the fused evaluator + ms_scan_affine.
//
// Host logic (coin, AIR, programs, wire format) comes from ministark_host.hpp and is CPU-tested.  This file only strings
// the ms_* calls together; it is compile- and link-checked in the build container and exercised on a GPU by
// tests/test_gpu_cpp_prover.py, which compares its proof bytes with the Python driver's.
#pragma once
#include <stdexcept>

#include "ministark_b200.h"
#include "ministark_examples.hpp"
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

// DEEP composition as a symbolic expression over LDE columns [base..., composition...] (ministark_b200/deep.py), with the
// terms grouped by their out-of-domain point:
//     sum_j a_j (P_j(x) - P_j(z_k)) / (x - z_k)  =  (sum_j a_j P_j(x) - K_k) / (x - z_k),      K_k = sum_j a_j P_j(z_k)
// one multiplication by a_j per column and one Fq x Fq product per distinct point (z^m and z*g^o per trace offset o).
// Hints, in order of first use: per group the alphas of its columns, the constant K_k and the point; then the two
// degree coefficients.
struct DeepKey { int kind; int64_t index; bool operator<(const DeepKey &o) const { return std::tie(kind, index) < std::tie(o.kind, o.index); } };
enum { DK_ZM, DK_KZM, DK_CALPHA, DK_ZPT, DK_KZ, DK_TALPHA, DK_DALPHA, DK_DBETA };
inline Expr deep_expression(Graph &g, const std::vector<std::pair<u64, int64_t>> &trace_arguments, u32 num_trace_cols,
                            u32 num_composition_cols, std::vector<DeepKey> &keys) {
    std::map<DeepKey, u64> index;
    auto H = [&](int kind, int64_t i) {
        DeepKey k{kind, i};
        if (!index.count(k)) { index[k] = keys.size(); keys.push_back(k); }
        return Hint(g, index[k]);
    };
    Expr x = X(g), one = Constant(g, 1), total;
    bool first = true;
    auto add_group = [&](const std::vector<std::pair<u64, DeepKey>> &members, const DeepKey &konst, const DeepKey &point) {
        if (members.empty()) return;
        Expr acc;
        bool f = true;
        for (const auto &m : members) {
            Expr term = Trace(g, m.first, 0) * H(m.second.kind, m.second.index);
            acc = f ? term : acc + term;
            f = false;
        }
        Expr term = (acc - H(konst.kind, konst.index)) * (one / (x - H(point.kind, point.index)));
        total = first ? term : total + term;
        first = false;
    };
    std::vector<std::pair<u64, DeepKey>> members;
    for (u32 j = 0; j < num_composition_cols; j++) members.push_back({num_trace_cols + j, DeepKey{DK_CALPHA, (int64_t)j}});
    add_group(members, DeepKey{DK_KZM, 0}, DeepKey{DK_ZM, 0});
    std::set<int64_t> offsets;
    for (const auto &ta : trace_arguments) offsets.insert(ta.second);
    for (int64_t off : offsets) {
        members.clear();
        for (size_t i = 0; i < trace_arguments.size(); i++)
            if (trace_arguments[i].second == off) members.push_back({trace_arguments[i].first, DeepKey{DK_TALPHA, (int64_t)i}});
        add_group(members, DeepKey{DK_KZ, off}, DeepKey{DK_ZPT, off});
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

    // builds the extension columns (num_extension_columns x n elements of Fq, column-major, device) from the resident base
    // trace and the challenges: Trace::build_extension_columns (src/trace.rs:27-34)
    using ExtensionBuilder = std::function<DeviceBuf(ms_ctx *, const u64 *base_dev, u64 n, const std::vector<Fq> &challenges)>;
    ms_ctx *context() const { return ctx; }

    // base_trace: num_base_columns x n Montgomery words, column-major, HOST memory.  public_inputs: handed to gen_hints;
    // public_inputs_bytes: their CanonicalSerialize form for the coin seed (empty: the Fq elements back to back, as for
    // examples/fib's single claimed value).
    Proof prove(const AirConfig &cfg, ProofOptions options, const u64 *base_trace, u64 n, const std::vector<Fq> &public_inputs,
                const Bytes &public_inputs_bytes = {}, const ExtensionBuilder &ext_builder = nullptr) {
        const u32 next = cfg.num_extension_columns;
        if (next && !ext_builder) throw std::runtime_error("this AIR has extension columns: pass an ExtensionBuilder");
        const int fq = cfg.fq_is_fp ? MS_FIELD_FP : MS_FIELD_FQ3, lanes = cfg.fq_is_fp ? 1 : 3;
        Air air(cfg, n, options);
        const unsigned log_n = air.log_n, beta = options.lde_blowup_factor, log_b = 31 - (unsigned)__builtin_clz(beta);
        const unsigned log_N = log_n + log_b;
        const u64 N = n << log_b, ce = air.ce_blowup_factor, M = n * ce;
        const unsigned log_ce = log_n + (63 - (unsigned)__builtin_clzll(ce));
        const u32 nbase = cfg.num_base_columns;
        const u64 GEN = to_mont(GENERATOR), ONE = to_mont(1);
        // gen_public_coin (examples/fib/main.rs:166-172)
        Bytes seed = public_inputs_bytes;
        if (seed.empty())
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
        coin.reseed_with_digest(proof.base_trace_commitment);
        std::vector<Fq> challenges;
        for (u64 i = 0; i < air.num_challenges(); i++) challenges.push_back(coin.draw());
        const std::vector<Fq> hints = cfg.gen_hints ? cfg.gen_hints(n, public_inputs, challenges) : std::vector<Fq>{};

        // ---- extension trace commitment (prover.rs:56-72)
        DeviceBuf ext_polys, ext_lde, ext_leaves, ext_nodes;
        if (next) {
            DeviceBuf ext = ext_builder(ctx, d_trace.words(), n, challenges);
            ext_polys = DeviceBuf(ctx, (size_t)next * n * lanes * 8);
            ext_lde = DeviceBuf(ctx, (size_t)next * N * lanes * 8);
            ext_leaves = DeviceBuf(ctx, N * 32);
            ext_nodes = DeviceBuf(ctx, N * 32);
            ck(ctx, ms_ntt_batch_to(ctx, fq, ext.p, n, ext_polys.p, n, next, log_n, MS_NTT_INVERSE, ONE), "extension interpolate");
            ck(ctx, ms_lde_batch(ctx, fq, ext_polys.p, n, ext_lde.p, N, next, log_n, log_b, GEN, 1), "extension lde");
            proof.has_extension = true;
            proof.extension_trace_commitment.resize(32);
            ck(ctx, ms_merkle_commit_sha256(ctx, fq, ext_lde.p, N, next, N, ext_leaves.p, ext_nodes.p, proof.extension_trace_commitment.data()),
               "extension commit");
            coin.reseed_with_digest(proof.extension_trace_commitment);
        }
        d_trace.release();

        // ---- constraint evaluation over the ce domain, read in place from the bit-reversed LDE prefix (prover.rs:75-108)
        std::vector<Fq> ccoefs;
        for (u64 i = 0; i < air.num_composition_constraint_coeffs(); i++) ccoefs.push_back(coin.draw());
        const Program prog = air.composition_program(nbase).bind(challenges, hints, ccoefs);
        DeviceBuf comp_evals(ctx, M * lanes * 8);
        ck(ctx, ms_eval_constraints(ctx, &prog.code[0][0], (unsigned)prog.code.size(), &prog.consts[0][0], (unsigned)prog.consts.size(),
                                    base_lde.p, N, nbase, next ? ext_lde.p : nullptr, N, next, fq, log_ce, GEN, 1, 0, comp_evals.p), "eval_constraints");

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
        std::vector<u64> ext_ood((size_t)next * offsets.size() * 3);
        if (next) ck(ctx, ms_poly_eval(ctx, fq, ext_polys.p, n, next, n, pts.data(), (unsigned)offsets.size(), ext_ood.data()), "ood (extension)");
        const u64 zm_w[3] = {to_mont(z_m.c[0]), to_mont(z_m.c[1]), to_mont(z_m.c[2])};
        ck(ctx, ms_poly_eval(ctx, fq, comp_polys, n, (unsigned)ce, n, zm_w, 1, comp_ood.data()), "ood (composition)");
        auto canon3 = [&](const u64 *w) {
            Fq v(from_mont(w[0]), from_mont(w[1]), from_mont(w[2]));
            if (lanes == 1 && (v.c[1] || v.c[2])) throw std::runtime_error("out-of-domain value left the base field although Fq = Fp");
            return v;
        };
        for (const auto &ta : trace_args) {
            const size_t k = std::find(offsets.begin(), offsets.end(), ta.second) - offsets.begin();
            if (ta.first < nbase) proof.execution_trace_ood_evals.push_back(canon3(&base_ood[(ta.first * offsets.size() + k) * 3]));
            else if (ta.first < nbase + next) proof.execution_trace_ood_evals.push_back(canon3(&ext_ood[((ta.first - nbase) * offsets.size() + k) * 3]));
            else throw std::runtime_error("trace argument names a column that does not exist");
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
        const Expr dexpr = deep_expression(dg, trace_args, nbase + next, (u32)ce, keys);
        std::vector<Fq> dhints;
        for (const DeepKey &k : keys) {
            switch (k.kind) {
                case DK_ZM: dhints.push_back(z_m); break;
                case DK_KZM: {      // sum_j alpha'_j H_j(z^m)
                    Fq acc;
                    for (u64 j = 0; j < ce; j++) acc = fq_add(acc, fq_mul(co_alphas[j], proof.composition_trace_ood_evals[j]));
                    dhints.push_back(acc);
                    break;
                }
                case DK_CALPHA: dhints.push_back(co_alphas[k.index]); break;
                case DK_ZPT: dhints.push_back(z_points[std::find(offsets.begin(), offsets.end(), k.index) - offsets.begin()]); break;
                case DK_KZ: {       // sum over the arguments with this offset of alpha * T(z g^o)
                    Fq acc;
                    for (size_t i = 0; i < trace_args.size(); i++)
                        if (trace_args[i].second == k.index) acc = fq_add(acc, fq_mul(ex_alphas[i], proof.execution_trace_ood_evals[i]));
                    dhints.push_back(acc);
                    break;
                }
                case DK_TALPHA: dhints.push_back(ex_alphas[k.index]); break;
                case DK_DALPHA: dhints.push_back(d_alpha); break;
                default: dhints.push_back(d_beta);
            }
        }
        const Program dprog = compile_program(dg, dexpr.id, nbase, 1, (int)log_N, /*batch_inverses=*/true).bind({}, dhints, {});
        std::vector<const void *> cols;
        std::vector<int> is_q;
        for (u32 c = 0; c < nbase; c++) { cols.push_back(base_lde.words() + (size_t)c * N); is_q.push_back(0); }
        for (u32 c = 0; c < next; c++) { cols.push_back(ext_lde.words() + (size_t)c * N * lanes); is_q.push_back(1); }
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
            if (ln < log_ff) throw std::runtime_error("FRI: the evaluation domain is smaller than the folding factor");
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
        if (next) {
            std::vector<u64> erow(positions.size() * next * lanes);
            ck(ctx, ms_gather_rows(ctx, fq, ext_lde.p, N, next, N, positions.data(), (unsigned)positions.size(), erow.data()), "extension rows");
            proof.trace_queries.extension_trace_values = canon_vec(erow, lanes);
            proof.trace_queries.has_extension = true;
            proof.trace_queries.extension_trace_proof = view_of(ext_leaves, ext_nodes, N, positions);
        }
        proof.trace_queries.base_trace_proof = view_of(base_leaves, base_nodes, N, positions);
        proof.trace_queries.composition_trace_proof = view_of(comp_leaves, comp_nodes, N, positions);
        return proof;
    }
};

// ------------------------------------------------------------------------------------------------ examples/brainfuck
namespace bf {

// public inputs of the brainfuck claim as ark-serialize writes them (main.rs:56-61): String, Vec<u8>, Vec<u8>
inline Bytes claim_bytes(const std::string &source, const Bytes &input, const Bytes &output) {
    Bytes o;
    put_u64_le(o, source.size());
    o.insert(o.end(), source.begin(), source.end());
    put_u64_le(o, input.size());
    o.insert(o.end(), input.begin(), input.end());
    put_u64_le(o, output.size());
    o.insert(o.end(), output.begin(), output.end());
    return o;
}

// The nine Fq3 extension columns (examples/brainfuck/trace.rs:108-279) on the device, as in
// ministark_b200/examples/brainfuck.py::_device_extension: every column is x_0 = init, x_(i+1) = x_i * a_i + b_i with
// per-row factors that are pointwise expressions of the base row (fused evaluator over the resident trace + eight 0/1
// helper columns derived from the integer rows), then one ms_scan_affine.  instr_initial / mem_initial: the two
// permutation start values (the reference draws them from ark_std::test_rng()).
inline DeviceBuf device_extension(ms_ctx *ctx, const VmTrace &t, const u64 *base_dev, const std::vector<Fq> &ch, const Fq &instr_initial,
                                  const Fq &mem_initial) {
    const u64 n = t.n;
    const unsigned log_n = 63 - (unsigned)__builtin_clzll(n);
    const u64 ONE = to_mont(1);
    // helper columns (Montgomery words)
    std::vector<u64> aux(8 * n);
    for (u64 r = 0; r < n; r++) {
        const u64 ci = t.at(CURR_INSTR, r), nxt_mv = t.at(MEM_VAL, (r + 1) % n), iip = t.at(I_IP, r);
        const bool same_ip = r > 0 && iip == t.at(I_IP, r - 1);
        const u64 v[8] = {ci != 0, ci == ',', ci == ',' ? nxt_mv : 0, ci == '.', ci == '.' ? nxt_mv : 0, t.at(M_DUMMY, r) == 0,
                          (u64)(t.at(I_CURR_INSTR, r) != 0 && same_ip), (u64)!same_ip};
        for (int k = 0; k < 8; k++) aux[(u64)k * n + r] = to_mont(v[k]);
    }
    DeviceBuf d_aux(ctx, aux.size() * 8);
    ck(ctx, ms_copy(ctx, d_aux.p, aux.data(), aux.size() * 8), "helper columns upload");
    std::vector<const void *> cols;
    std::vector<int> is_q(25, 0);
    for (u32 c = 0; c < 17; c++) cols.push_back(base_dev + (u64)c * n);
    for (u32 k = 0; k < 8; k++) cols.push_back(d_aux.words() + (u64)k * n);
    auto evaluate = [&](Graph &g, const Expr &e) {
        DeviceBuf out(ctx, n * 24);
        const Program p = compile_program(g, e.id, 25, 1, (int)log_n).bind(ch, {}, {});
        ck(ctx, ms_eval_constraints_ptrs(ctx, &p.code[0][0], (unsigned)p.code.size(), &p.consts[0][0], (unsigned)p.consts.size(), cols.data(),
                                         is_q.data(), 25, MS_FIELD_FQ3, log_n, ONE, 0, 0, out.p), "extension factors");
        return out;
    };
    Graph g;
    auto T = [&](u32 c) { return Trace(g, c, 0); };
    auto AUX = [&](u32 k) { return Trace(g, 17 + k, 0); };
    auto CH = [&](u64 i) { return Challenge(g, i); };
    const Expr one = Constant(g, 1);
    auto instr_fp = [&](u32 ip, u32 c, u32 nx) { return CH(CH_ALPHA) - CH(CH_A) * T(ip) - CH(CH_B) * T(c) - CH(CH_C) * T(nx); };
    auto mem_fp = [&](u32 cy, u32 mp, u32 v) { return CH(CH_BETA) - CH(CH_D) * T(cy) - CH(CH_E) * T(mp) - CH(CH_F) * T(v); };
    auto gated = [&](const Expr &mask, const Expr &factor) { return one + mask * (factor - one); };   // factor where mask = 1, else 1
    DeviceBuf ext(ctx, (size_t)9 * n * 24);
    auto words3 = [](const Fq &v) { return std::array<u64, 3>{to_mont(v.c[0]), to_mont(v.c[1]), to_mont(v.c[2])}; };
    const std::array<u64, 3> zero3 = {0, 0, 0}, ii = words3(instr_initial), mi = words3(mem_initial);
    auto scan = [&](u32 k, const std::array<u64, 3> &init, const void *a, const u64 *a_const, const void *b, int b_field, int inclusive) {
        ck(ctx, ms_scan_affine(ctx, MS_FIELD_FQ3, a, MS_FIELD_FQ3, a_const, b, b_field, n, init.data(), inclusive, ext.words() + (u64)k * n * 3), "scan");
    };
    { DeviceBuf a = evaluate(g, gated(AUX(0), instr_fp(IP, CURR_INSTR, NEXT_INSTR))); scan(0, ii, a.p, nullptr, nullptr, 1, 0); }
    { DeviceBuf a = evaluate(g, gated(AUX(0), mem_fp(CYCLE, MP, MEM_VAL))); scan(1, mi, a.p, nullptr, nullptr, 1, 0); }
    { DeviceBuf a = evaluate(g, gated(AUX(1), CH(CH_GAMMA))); scan(2, zero3, a.p, nullptr, d_aux.words() + 2 * n, MS_FIELD_FP, 0); }
    { DeviceBuf a = evaluate(g, gated(AUX(3), CH(CH_DELTA))); scan(3, zero3, a.p, nullptr, d_aux.words() + 4 * n, MS_FIELD_FP, 0); }
    { DeviceBuf a = evaluate(g, gated(AUX(5), mem_fp(M_CYCLE, M_MP, M_MEM_VAL))); scan(4, mi, a.p, nullptr, nullptr, 1, 0); }
    { DeviceBuf a = evaluate(g, gated(AUX(6), instr_fp(I_IP, I_CURR_INSTR, I_NEXT_INSTR))); scan(5, ii, a.p, nullptr, nullptr, 1, 1); }
    {
        DeviceBuf a = evaluate(g, gated(AUX(7), CH(CH_ETA)));
        DeviceBuf b = evaluate(g, AUX(7) * (CH(CH_A) * T(I_IP) + CH(CH_B) * T(I_CURR_INSTR) + CH(CH_C) * T(I_NEXT_INSTR)));
        scan(6, zero3, a.p, nullptr, b.p, MS_FIELD_FQ3, 1);
    }
    const std::array<u64, 3> gamma = words3(ch.at(CH_GAMMA)), delta = words3(ch.at(CH_DELTA));
    scan(7, zero3, nullptr, gamma.data(), base_dev + (u64)IN_VALUE * n, MS_FIELD_FP, 1);
    scan(8, zero3, nullptr, delta.data(), base_dev + (u64)OUT_VALUE * n, MS_FIELD_FP, 1);
    ck(ctx, ms_ctx_sync(ctx), "extension sync");       // the factor buffers above are freed when their scopes end
    return ext;
}

}  // namespace bf
}  // namespace mshost
