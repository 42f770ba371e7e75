// cpu_prover_main.cpp — CPU ORACLE.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// The C++ `default_prove` of include/ministark_prover.hpp linked against the CPU build of the C ABI (cpu_abi.c): a prover
// compiled end to end, run on the host cores by `bench.py --impl reference` as the compiled CPU baseline of the full
// prove (examples/fib, the workload of the GPU arm's `full_prove`).  Same formulation as the GPU driver (evaluation-form
// DEEP, per-coset FRI fold), so its kind is "port"; the reference's own formulation is oracle/stark_oracle.cpu_prove,
// whose proof bytes it must reproduce (tests/test_cpp_cpu_abi.py).
// prints one JSON line: {"seconds": prove time (trace generation excluded), "proof_bytes": …, "verified": true, …}
#include <chrono>
#include <cstdio>
#include <cstring>

#include "ministark_prover.hpp"
#include "ministark_verifier.hpp"

extern "C" int orc_num_threads(void);

using namespace mshost;

static void report(const Bytes &bytes, double prove_s, double verify_s, double trace_s, u64 rows, bool hex) {
    printf("{\"seconds\": %.6f, \"verify_seconds\": %.6f, \"trace_seconds\": %.6f, \"rows\": %llu, \"proof_bytes\": %zu, \"verified\": true, "
           "\"threads\": %d", prove_s, verify_s, trace_s, (unsigned long long)rows, bytes.size(), orc_num_threads());
    if (hex) {
        printf(", \"proof_hex\": \"");
        for (u8 b : bytes) printf("%02x", b);
        printf("\"");
    }
    printf("}\n");
}
static double since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

//   cpu_prover <log_rows> <nq> <blowup> <grind> <ff> <maxrem> [--hex]                                  examples/fib
//   cpu_prover bf <a> <b> <c> <nq> <blowup> <grind> <ff> <maxrem> <6 canonical integers> [--hex]       examples/brainfuck,
//       program cycle_burner(a, b, c) (40 40 60: 2^20 rows), the two permutation start values as in prover_test.cpp
int main(int argc, char **argv) {
    const bool bf_mode = argc > 1 && !strcmp(argv[1], "bf");
    if (argc < (bf_mode ? 16 : 7)) { fprintf(stderr, "usage: cpu_prover log_rows nq blowup grind ff maxrem [--hex] | cpu_prover bf a b c nq blowup grind ff maxrem i0 i1 i2 m0 m1 m2 [--hex]\n"); return 2; }
    char **o = argv + (bf_mode ? 5 : 2);
    const ProofOptions opts{(u8)atoi(o[0]), (u8)atoi(o[1]), (u8)atoi(o[2]), (u8)atoi(o[3]), (u8)atoi(o[4])};
    const bool hex = !strcmp(argv[argc - 1], "--hex");
    try {
        GpuProver prover(0);                               // the "device" is the host: libms_cpu_abi.so
        if (!bf_mode) {
            const u64 n = (u64)1 << atoi(argv[1]);
            std::vector<u64> trace;
            auto t0 = std::chrono::steady_clock::now();
            const u64 last = fib_gen_trace(n, trace);          // examples/fib/main.rs:175-222
            const double trace_s = since(t0);
            t0 = std::chrono::steady_clock::now();
            const Proof proof = prover.prove(fib_air_config(), opts, trace.data(), n, {Fq(last)});
            const double prove_s = since(t0);
            const Bytes bytes = proof.to_bytes(1);
            t0 = std::chrono::steady_clock::now();
            verify(fib_air_config(), bytes, {Fq(last)}, {}, 10);
            report(bytes, prove_s, since(t0), trace_s, n, hex);
        } else {
            const std::string src = bf::cycle_burner((unsigned)atoi(argv[2]), (unsigned)atoi(argv[3]), (unsigned)atoi(argv[4]));
            auto t0 = std::chrono::steady_clock::now();
            const bf::VmTrace t = bf::simulate(src);           // examples/brainfuck/vm.rs
            std::vector<u64> words(t.base.size());
            for (size_t i = 0; i < words.size(); i++) words[i] = to_mont(t.base[i]);
            const double trace_s = since(t0);
            const Fq ii(strtoull(argv[10], nullptr, 10), strtoull(argv[11], nullptr, 10), strtoull(argv[12], nullptr, 10));
            const Fq mi(strtoull(argv[13], nullptr, 10), strtoull(argv[14], nullptr, 10), strtoull(argv[15], nullptr, 10));
            const AirConfig cfg = bf::air_config(src, {}, t.output);
            const Bytes claim = bf::claim_bytes(src, {}, t.output);
            t0 = std::chrono::steady_clock::now();
            const Proof proof = prover.prove(cfg, opts, words.data(), t.n, {}, claim, [&](ms_ctx *ctx, const u64 *base_dev, u64, const std::vector<Fq> &ch) {
                return bf::device_extension(ctx, t, base_dev, ch, ii, mi);
            });
            const double prove_s = since(t0);
            const Bytes bytes = proof.to_bytes(3);
            t0 = std::chrono::steady_clock::now();
            verify(cfg, bytes, {}, claim, 10);
            report(bytes, prove_s, since(t0), trace_s, t.n, hex);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "cpu_prover: %s\n", e.what());
        return 1;
    }
    return 0;
}
