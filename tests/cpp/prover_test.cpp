// GPU driver for include/ministark_prover.hpp: proves examples/fib with the C++ host layer + the C ABI and prints
//   <claimed value> <proof bytes as hex>
// usage: prover_test <log_rows> <num_queries> <blowup> <grinding> <folding> <max_remainder_coeffs>
#include <cstdio>
#include <iostream>

#include "ministark_prover.hpp"

using namespace mshost;

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: prover_test log_rows nq blowup grind ff maxrem\n"); return 2; }
    const unsigned log_rows = (unsigned)atoi(argv[1]);
    const ProofOptions opts{(u8)atoi(argv[2]), (u8)atoi(argv[3]), (u8)atoi(argv[4]), (u8)atoi(argv[5]), (u8)atoi(argv[6])};
    const u64 n = (u64)1 << log_rows;
    // gen_trace (examples/fib/main.rs:175-222): v_k = v_(k-2) * v_(k-1), 8 consecutive values per row; Montgomery words
    std::vector<u64> trace(8 * n);
    u64 v[8] = {1, 2, 0, 0, 0, 0, 0, 0};
    for (int i = 2; i < 8; i++) v[i] = mulm(v[i - 2], v[i - 1]);
    u64 last = 0;
    for (u64 r = 0; r < n; r++) {
        for (int c = 0; c < 8; c++) trace[(u64)c * n + r] = to_mont(v[c]);
        last = v[7];
        u64 w[8];
        w[0] = mulm(v[6], v[7]);
        w[1] = mulm(v[7], w[0]);
        for (int i = 2; i < 8; i++) w[i] = mulm(w[i - 2], w[i - 1]);
        for (int i = 0; i < 8; i++) v[i] = w[i];
    }
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
