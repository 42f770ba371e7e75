// ministark_gpu.hpp — C++ mirror of the Rust item set the reference's main crate imports from
// `ministark-gpu` (gpu/src/prelude.rs:1-19) plus the Matrix / Merkle seams (src/matrix.rs,
// src/merkle.rs), written over the C ABI of include/ministark_b200.h.  Header only.
//
// Same names, argument meaning and error behaviour as the reference: where the Rust code panics
// (assert!/unwrap, gpu/src/stage.rs:55-75, gpu/src/plan.rs:255-257) these throw std::runtime_error.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ministark_b200.h"

namespace ministark_gpu {

using u64 = uint64_t;
constexpr u64 FP_ONE = 4294967295ull;          // Fp::ONE, Montgomery (felt_u64.h.metal:118)
constexpr u64 FP_GENERATOR = 30064771065ull;   // Fp::GENERATOR = 7, Montgomery

// GpuField (gpu/src/lib.rs:20-38): a field that has kernels; FIELD_ID selects them where the
// reference uses field_name() strings (gpu/src/fields.rs:58-60,214-216).
struct Fp  { static constexpr int FIELD_ID = MS_FIELD_FP;  static constexpr int WORDS = 1; };
struct Fq3 { static constexpr int FIELD_ID = MS_FIELD_FQ3; static constexpr int WORDS = 3; };

// Planner / get_planner() (gpu/src/plan.rs:327-350,465-469)
class Planner {
   public:
    explicit Planner(int device = 0) {
        if (ms_ctx_create(device, &ctx_) != MS_OK) throw std::runtime_error("no device found");  // plan.rs:467
    }
    ~Planner() { if (ctx_) ms_ctx_destroy(ctx_); }
    Planner(const Planner &) = delete;
    Planner &operator=(const Planner &) = delete;
    ms_ctx *ctx() const { return ctx_; }
    void check(int rc) const {
        if (rc != MS_OK) throw std::runtime_error(std::string("ministark_gpu: ") + ms_last_error(ctx_));
    }

   private:
    ms_ctx *ctx_ = nullptr;
};
inline Planner &get_planner() {
    static Planner p;   // lazily created process-global, like once_cell::Lazy<Planner>
    return p;
}

// Radix2EvaluationDomain<F::FftField>: size = 2^log_size, coset offset (Montgomery word)
struct Radix2EvaluationDomain {
    unsigned log_size;
    u64 offset = FP_ONE;
    size_t size() const { return size_t(1) << log_size; }
};

template <class F, int DIRECTION>
class FftBase {
   public:
    static constexpr size_t MIN_SIZE = 2048;   // gpu/src/plan.rs:246,292 (interface parity; not enforced)
    explicit FftBase(Radix2EvaluationDomain domain, Planner &pl = get_planner()) : pl_(pl), n_(domain.size()) {
        pl_.check(ms_ntt_plan_create(pl_.ctx(), F::FIELD_ID, domain.log_size, DIRECTION, domain.offset, &plan_));
    }
    ~FftBase() { if (plan_) ms_ntt_plan_destroy(plan_); }
    // encode(&mut [F]): in-place transform of a caller-owned slice of exactly domain.size() elements
    void encode(u64 *buffer, size_t len) {
        if (len != n_) throw std::runtime_error("assertion failed: encoder.n == buffer.len()");  // plan.rs:257,303
        pl_.check(ms_ntt_encode(plan_, buffer));
    }
    // execute(self): commit + wait_until_completed (plan.rs:229-232)
    void execute() { pl_.check(ms_ntt_execute(plan_)); }

   private:
    Planner &pl_;
    size_t n_;
    ms_ntt_plan *plan_ = nullptr;
};
template <class F> using GpuFft = FftBase<F, MS_NTT_FORWARD>;    // gpu/src/plan.rs:236-279
template <class F> using GpuIfft = FftBase<F, MS_NTT_INVERSE>;   // gpu/src/plan.rs:282-325

// Stage wrappers (gpu/src/stage.rs): every *Stage::encode maps onto ms_pointwise / ms_pointwise_const.
template <class L, class R = L>
struct MulAssignStage {   // lhs[i] *= rhs[(i + shift) % n]   (stage.rs:175-232)
    size_t n;
    void encode(Planner &pl, u64 *lhs, const u64 *rhs, size_t shift = 0) const {
        pl.check(ms_pointwise(pl.ctx(), MS_OP_MUL, L::FIELD_ID, lhs, L::FIELD_ID, lhs, R::FIELD_ID, rhs, n, shift, 0));
    }
};
template <class L, class R = L>
struct AddAssignStage {   // lhs[i] += rhs[(i + shift) % n]   (stage.rs:396-452; used by sum_columns_gpu)
    size_t n;
    void encode(Planner &pl, u64 *lhs, const u64 *rhs, size_t shift = 0) const {
        pl.check(ms_pointwise(pl.ctx(), MS_OP_ADD, L::FIELD_ID, lhs, L::FIELD_ID, lhs, R::FIELD_ID, rhs, n, shift, 0));
    }
};
template <class F>
struct InverseInPlaceStage {   // stage.rs:805-844
    size_t n;
    void encode(Planner &pl, u64 *buf) const {
        pl.check(ms_pointwise(pl.ctx(), MS_OP_INV, F::FIELD_ID, buf, F::FIELD_ID, buf, MS_FIELD_FP, nullptr, n, 0, 0));
    }
};
template <class F>
struct ExpInPlaceStage {       // stage.rs:993-1032
    size_t n;
    void encode(Planner &pl, u64 *buf, uint32_t exponent) const {
        pl.check(ms_pointwise(pl.ctx(), MS_OP_EXP, F::FIELD_ID, buf, F::FIELD_ID, buf, MS_FIELD_FP, nullptr, n, 0, exponent));
    }
};
template <class F>
struct BitReverseGpuStage {    // stage.rs:280-332
    unsigned log_n;
    void encode(Planner &pl, u64 *buf) const { pl.check(ms_bit_reverse(pl.ctx(), F::FIELD_ID, buf, size_t(1) << log_n, 1, log_n)); }
};

// Matrix<F> (src/matrix.rs:26): columns of equal length, column-major, contiguous here.
template <class F>
class Matrix {
   public:
    Matrix(size_t num_cols, size_t num_rows) : cols_(num_cols), rows_(num_rows), data_(num_cols * num_rows * F::WORDS) {}
    size_t num_cols() const { return cols_; }
    size_t num_rows() const { return rows_; }
    u64 *column(size_t c) { return data_.data() + c * rows_ * F::WORDS; }
    u64 *data() { return data_.data(); }

    // interpolate (src/matrix.rs:157-163)
    Matrix interpolate(Radix2EvaluationDomain d, Planner &pl = get_planner()) const {
        Matrix out(*this);
        pl.check(ms_ntt_batch(pl.ctx(), F::FIELD_ID, out.data(), rows_, (unsigned)cols_, d.log_size, MS_NTT_INVERSE, d.offset));
        return out;
    }
    // evaluate / bit_reversed_evaluate (src/matrix.rs:237-251)
    Matrix evaluate(Radix2EvaluationDomain d, bool bit_reversed, Planner &pl = get_planner()) const {
        unsigned log_rows = 0;
        while ((size_t(1) << log_rows) < rows_) log_rows++;
        if ((size_t(1) << log_rows) != rows_ || d.log_size < log_rows) throw std::runtime_error("bad domain");
        Matrix out(cols_, d.size());
        pl.check(ms_lde_batch(pl.ctx(), F::FIELD_ID, data_.data(), rows_, out.data(), d.size(), (unsigned)cols_, log_rows,
                              d.log_size - log_rows, d.offset, bit_reversed ? 1 : 0));
        return out;
    }
    // sum_columns (src/matrix.rs:322-394)
    Matrix sum_columns(Planner &pl = get_planner()) const {
        Matrix out(1, rows_);
        pl.check(ms_sum_columns(pl.ctx(), F::FIELD_ID, data_.data(), rows_, (unsigned)cols_, rows_, out.data()));
        return out;
    }

   private:
    size_t cols_, rows_;
    std::vector<u64> data_;
};

// MatrixMerkleTreeImpl<Sha256HashFn>::from_matrix (src/merkle.rs:359-361): nodes in heap layout.
struct MatrixMerkleTree {
    std::vector<uint8_t> leaves, nodes;
    template <class F>
    static MatrixMerkleTree from_matrix(Matrix<F> &m, Planner &pl = get_planner()) {
        MatrixMerkleTree t;
        t.leaves.resize(m.num_rows() * 32);
        t.nodes.resize(m.num_rows() * 32);
        uint8_t root[32];
        pl.check(ms_merkle_commit_sha256(pl.ctx(), F::FIELD_ID, m.data(), m.num_rows(), (unsigned)m.num_cols(), m.num_rows(),
                                         t.leaves.data(), t.nodes.data(), root));
        return t;
    }
    const uint8_t *root() const { return nodes.data() + 32; }   // nodes[1] (src/merkle.rs:143-145)
};

}  // namespace ministark_gpu
