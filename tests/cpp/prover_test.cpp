// GPU driver for include/ministark_prover.hpp: proves an example with the C++ host layer + the C ABI and prints
//   <claim> <proof bytes as hex>
//   prover_test fib <log_rows> <num_queries> <blowup> <grinding> <folding> <max_remainder_coeffs>
//   prover_test bf hello|burner:<a>:<b>:<c> <nq> <blowup> <grind> <ff> <maxrem> <6 canonical integers: instr_initial, mem_initial>
#include <cstdio>
#include <iostream>

#include "ministark_prover.hpp"
#include "ministark_verifier.hpp"

using namespace mshost;

static std::string hex(const Bytes &b) {
    static const char *d = "0123456789abcdef";
    std::string s;
    for (u8 c : b) { s.push_back(d[c >> 4]); s.push_back(d[c & 15]); }
    return s;
}

int main(int argc, char **argv) {
    if (argc < 8) { fprintf(stderr, "usage: prover_test fib|bf ... (see the header of this file)\n"); return 2; }
    const std::string kind = argv[1];
    const ProofOptions opts{(u8)atoi(argv[3]), (u8)atoi(argv[4]), (u8)atoi(argv[5]), (u8)atoi(argv[6]), (u8)atoi(argv[7])};
    try {
        GpuProver prover(0);
        if (kind == "fib") {
            const u64 n = (u64)1 << atoi(argv[2]);
            std::vector<u64> trace;
            const u64 last = fib_gen_trace(n, trace);          // examples/fib/main.rs:175-222
            const Proof proof = prover.prove(fib_air_config(), opts, trace.data(), n, {Fq(last)});
            verify(fib_air_config(), proof.to_bytes(1), {Fq(last)}, {}, 10);        // the C++ verifier accepts its own prover's proof
            std::cout << last << " " << hex(proof.to_bytes(1)) << "\n";
        } else {
            if (argc < 14) { fprintf(stderr, "bf needs the two initial values\n"); return 2; }
            const std::string which = argv[2];
            std::string src = "++++++++++[>+++++++>++++++++++>+++>+<<<<-]>++.>+.+++++++..+++.>++.<<+++++++++++++++.>.+++.------.--------.";
            unsigned a, b, c;
            if (sscanf(which.c_str(), "burner:%u:%u:%u", &a, &b, &c) == 3) src = bf::cycle_burner(a, b, c);
            const bf::VmTrace t = bf::simulate(src);
            std::vector<u64> words(t.base.size());
            for (size_t i = 0; i < words.size(); i++) words[i] = to_mont(t.base[i]);
            const Fq ii(strtoull(argv[8], nullptr, 10), strtoull(argv[9], nullptr, 10), strtoull(argv[10], nullptr, 10));
            const Fq mi(strtoull(argv[11], nullptr, 10), strtoull(argv[12], nullptr, 10), strtoull(argv[13], nullptr, 10));
            const Proof proof = prover.prove(bf::air_config(src, {}, t.output), opts, words.data(), t.n, {}, bf::claim_bytes(src, {}, t.output),
                                             [&](ms_ctx *ctx, const u64 *base_dev, u64, const std::vector<Fq> &ch) {
                                                 return bf::device_extension(ctx, t, base_dev, ch, ii, mi);
                                             });
            verify(bf::air_config(src, {}, t.output), proof.to_bytes(3), {}, bf::claim_bytes(src, {}, t.output), 10);
            std::cout << "out:" << hex(t.output) << " " << hex(proof.to_bytes(3)) << "\n";
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "prover_test: %s\n", e.what());
        return 1;
    }
    return 0;
}
