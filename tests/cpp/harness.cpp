// C++ harness over include/ministark_gpu.hpp: the same call sequence the reference's Rust makes
// (src/prover.rs:46-55): interpolate -> bit_reversed_evaluate -> MatrixMerkleTree::from_matrix, plus a
// GpuFft encode/execute round trip.  usage: harness <in.bin> <out.bin> log_n ncols log_blowup
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ministark_gpu.hpp"

using namespace ministark_gpu;

int main(int argc, char **argv) {
    if (argc != 6) return 2;
    const unsigned log_n = atoi(argv[3]), ncols = atoi(argv[4]), log_b = atoi(argv[5]);
    const size_t n = size_t(1) << log_n;
    Matrix<Fp> trace(ncols, n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(trace.data(), 8, ncols * n, f) != ncols * n) return 3;
    fclose(f);
    try {
        Matrix<Fp> polys = trace.interpolate(Radix2EvaluationDomain{log_n});
        Matrix<Fp> lde = polys.evaluate(Radix2EvaluationDomain{log_n + log_b, FP_GENERATOR}, /*bit_reversed=*/true);
        MatrixMerkleTree tree = MatrixMerkleTree::from_matrix(lde);
        // GpuIfft / GpuFft on column 0, in place on a host slice
        std::vector<u64> col(trace.column(0), trace.column(0) + n);
        GpuIfft<Fp> ifft(Radix2EvaluationDomain{log_n});
        ifft.encode(col.data(), col.size());
        ifft.execute();
        if (memcmp(col.data(), polys.column(0), n * 8) != 0) return 4;
        GpuFft<Fp> fft(Radix2EvaluationDomain{log_n});
        fft.encode(col.data(), col.size());
        fft.execute();
        if (memcmp(col.data(), trace.column(0), n * 8) != 0) return 5;
        bool threw = false;
        try { fft.encode(col.data(), col.size() - 1); } catch (const std::runtime_error &) { threw = true; }
        if (!threw) return 6;   // the reference asserts encoder.n == buffer.len()
        f = fopen(argv[2], "wb");
        fwrite(tree.root(), 1, 32, f);
        fwrite(polys.data(), 8, ncols * n, f);
        fclose(f);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
