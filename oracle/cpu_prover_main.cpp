// cpu_prover_main.cpp — CPU ORACLE.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// The C++ `default_prove` of include/ministark_prover.hpp linked against the CPU build of the C ABI (cpu_abi.c): a prover
// compiled end to end, run on the host cores by `bench.py --impl reference` as the compiled CPU baseline of the full
// prove (examples/fib, the workload of the GPU arm's `full_prove`).  Same formulation as the GPU driver (evaluation-form
// DEEP, per-coset FRI fold), so its kind is "port"; the reference's own formulation is oracle/stark_oracle.cpu_prove,
// whose proof bytes it must reproduce (tests/test_cpp_cpu_abi.py).
//   cpu_prover <log_rows> <num_queries> <blowup> <grinding> <folding> <max_remainder_coeffs> [--hex]
// prints one JSON line: {"seconds": prove time (trace generation excluded), "proof_bytes": …, "verified": true, …}
#include <chrono>
#include <cstdio>
#include <cstring>

#include "ministark_prover.hpp"
#include "ministark_verifier.hpp"

extern "C" int orc_num_threads(void);

using namespace mshost;

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: cpu_prover log_rows nq blowup grind ff maxrem [--hex]\n"); return 2; }
    const ProofOptions opts{(u8)atoi(argv[2]), (u8)atoi(argv[3]), (u8)atoi(argv[4]), (u8)atoi(argv[5]), (u8)atoi(argv[6])};
    const bool hex = argc > 7 && !strcmp(argv[7], "--hex");
    try {
        const u64 n = (u64)1 << atoi(argv[1]);
        std::vector<u64> trace;
        const u64 last = fib_gen_trace(n, trace);          // examples/fib/main.rs:175-222
        GpuProver prover(0);                               // the "device" is the host: libms_cpu_abi.so
        const auto t0 = std::chrono::steady_clock::now();
        const Proof proof = prover.prove(fib_air_config(), opts, trace.data(), n, {Fq(last)});
        const double prove_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const Bytes bytes = proof.to_bytes(1);
        const auto t1 = std::chrono::steady_clock::now();
        verify(fib_air_config(), bytes, {Fq(last)}, {}, 10);
        const double verify_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        printf("{\"seconds\": %.6f, \"verify_seconds\": %.6f, \"proof_bytes\": %zu, \"verified\": true, \"threads\": %d, \"claim\": %llu",
               prove_s, verify_s, bytes.size(), orc_num_threads(), (unsigned long long)last);
        if (hex) {
            printf(", \"proof_hex\": \"");
            for (u8 b : bytes) printf("%02x", b);
            printf("\"");
        }
        printf("}\n");
    } catch (const std::exception &e) {
        fprintf(stderr, "cpu_prover: %s\n", e.what());
        return 1;
    }
    return 0;
}
