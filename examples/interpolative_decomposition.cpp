// examples/interpolative_decomposition.rs of the reference, line for line, on the C++ host mirror
// (include/rusty_compression_b200.hpp).  Build:
//   g++ -std=c++17 -Iinclude examples/interpolative_decomposition.cpp -Lrusty_compression_b200 -lrc_b200
#include <cstdio>
#include <cstdlib>

#include "rusty_compression_b200.hpp"

using namespace rcb200;

int main(int argc, char** argv) {
    // We initialize a random number generator (here: the seed of the device Philox generator).
    const uint64_t seed = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 0;
    Context ctx(0);

    // The dimension of the matrix for which we want to compute a low-rank approximation, and the compression rank.
    const size_t rows = 500, cols = 100, k = 20;

    // Generate a random matrix with singular values logarithmically distributed between 1 and 1E-10.
    auto mat = Matrix<double>::random_approximate_low_rank_matrix(ctx, rows, cols, 1.0, 1e-10, seed);

    // Compute the pivoted QR decomposition of the matrix.
    auto qr = QR<double>::compute_from(mat);

    // Compress it to only include the k most significant basis vectors of the range.
    auto qr_compressed = qr.compress(CompressionType::RANK(k));

    // From the compressed representation compute the column interpolative decomposition ...
    auto col_int_decomp = qr_compressed.column_id();

    // ... and a two sided interpolative decomposition.
    auto two_sided_int_decomp = col_int_decomp.two_sided_id();

    // Multiply the factors back (debugging / non-probabilistic error computation only) and compare.
    auto mat_approx = two_sided_int_decomp.to_mat();
    const double rel_diff = rel_diff_fro(mat, mat_approx);
    std::printf("The relative difference of the compressed and original matrix is %1.2E\n", rel_diff);

    // quirk Q3: tolerance compression errors when no diagonal entry falls below the tolerance
    bool raised = false;
    try {
        qr_compressed.compress(CompressionType::ADAPTIVE(1e-30));
    } catch (const CompressionError&) {
        raised = true;
    }
    std::printf("compress(ADAPTIVE(1e-30)) raised CompressionError: %s\n", raised ? "yes" : "no");
    return (rel_diff > 0.0 && rel_diff < 0.2 && raised && two_sided_int_decomp.rank() == k &&
            two_sided_int_decomp.get_col_ind().size() == cols && two_sided_int_decomp.get_row_ind().size() == rows) ? 0 : 2;
}
