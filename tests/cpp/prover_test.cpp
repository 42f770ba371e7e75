// GPU driver for include/ministark_prover.hpp: proves examples/fib with the C++ host layer + the C ABI and prints
//   <claimed value> <proof bytes as hex>
// usage: prover_test <log_rows> <num_queries> <blowup> <grinding> <folding> <max_remainder_coeffs>
#include <cstdio>
#include <iostream>

#include "ministark_examples.hpp"
#include "ministark_prover.hpp"

using namespace mshost;

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: prover_test log_rows nq blowup grind ff maxrem\n"); return 2; }
    const unsigned log_rows = (unsigned)atoi(argv[1]);
    const ProofOptions opts{(u8)atoi(argv[2]), (u8)atoi(argv[3]), (u8)atoi(argv[4]), (u8)atoi(argv[5]), (u8)atoi(argv[6])};
    const u64 n = (u64)1 << log_rows;
    std::vector<u64> trace;
    const u64 last = fib_gen_trace(n, trace);          // examples/fib/main.rs:175-222
    try {
        GpuProver prover(0);
        const Proof proof = prover.prove(fib_air_config(), opts, trace.data(), n, {Fq(last)});
        static const char *d = "0123456789abcdef";
        std::string s;
        for (u8 c : proof.to_bytes(1)) { s.push_back(d[c >> 4]); s.push_back(d[c & 15]); }
        std::cout << last << " " << s << "\n";
    } catch (const std::exception &e) {
        fprintf(stderr, "prover_test: %s\n", e.what());
        return 1;
    }
    return 0;
}
