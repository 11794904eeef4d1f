// The reference's unit tests that go through the public surface, restated on the C++ host mirror
// (include/rusty_compression_b200.hpp) with the reference's shapes and thresholds:
//   pivoted_qr_tests / pivoted_lq_tests          src/pivoted_qr.rs:198-316   (100 x 50 and 50 x 100; rel. error < tol, Q^H Q = I)
//   qr_compression_by_rank / _by_tol tests       src/qr.rs:427-560
//   id_compression_tests (column / row ID)       src/col_interp_decomp.rs:176-243, src/row_interp_decomp.rs:176-237
//   svd compression tests                        src/svd.rs:200-290
//   permutation known answers                    src/permutation.rs:192-239
// Prints one line per check and exits non-zero on the first failure.  Needs a B200 (no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "rusty_compression_b200.hpp"

using namespace rcb200;

static int failures = 0;
static void expect(bool ok, const std::string& what) {
    std::printf("%s  %s\n", ok ? "ok  " : "FAIL", what.c_str());
    if (!ok) ++failures;
}

template <class A>
static Matrix<A> eye(const Context& ctx, size_t n) {
    std::vector<A> h(n * n, A(0));
    for (size_t i = 0; i < n; ++i) h[i * n + i] = A(1);
    return Matrix<A>::from_host(ctx, h, n, n);
}

template <class A>
static Matrix<A> x_for_apply(const Context& ctx, size_t cols, A) { return Matrix<A>::random_gaussian(ctx, cols, 3, 5); }

template <class A>
static void factorization_tests(const Context& ctx, const char* name, double tol, size_t rows, size_t cols) {
    const std::string tag = std::string(name) + " " + std::to_string(rows) + "x" + std::to_string(cols) + ": ";
    auto mat = Matrix<A>::random_approximate_low_rank_matrix(ctx, rows, cols, 1.0, 1e-10, 17 + rows);
    const size_t kmin = rows < cols ? rows : cols;

    auto qr = QR<A>::compute_from(mat);                                   // src/pivoted_qr.rs:198-246
    expect(qr.get_q().nrows() == rows && qr.get_q().ncols() == kmin && qr.get_r().nrows() == kmin && qr.get_r().ncols() == cols,
           tag + "pivoted_qr shapes");
    expect(rel_diff_fro(qr.to_mat(), mat) < tol, tag + "pivoted_qr reconstruction");
    expect(rel_diff_fro(qr.get_q().conj_matmat(qr.get_q()), eye<A>(ctx, kmin)) < tol, tag + "Q^H Q = I");
    expect(qr.get_ind().size() == cols, tag + "ind has full length (quirk Q8)");

    auto lq = LQ<A>::compute_from(mat);                                   // src/pivoted_qr.rs:248-294
    expect(rel_diff_fro(lq.to_mat(), mat) < tol, tag + "pivoted_lq reconstruction");

    auto qr30 = qr.compress(CompressionType::RANK(30));                   // src/qr.rs:427-462
    expect(qr30.rank() == 30 && qr30.get_q().ncols() == 30 && qr30.get_r().nrows() == 30, tag + "compress_qr_rank shapes");
    auto qrt = qr.compress(CompressionType::ADAPTIVE(1e-4));              // src/qr.rs:464-494
    expect(rel_diff_fro(qrt.to_mat(), mat) < 5e-4 * std::sqrt((double)kmin), tag + "compress_qr_tolerance error");

    auto cid = qrt.column_id();                                           // src/col_interp_decomp.rs:176-230
    expect(rel_diff_fro(cid.to_mat(), mat) < 5e-4 * std::sqrt((double)kmin), tag + "column_id error");
    expect(cid.get_col_ind().size() == cols && cid.rank() == qrt.rank(), tag + "column_id shapes");
    auto tid = cid.two_sided_id();
    expect(rel_diff_fro(tid.to_mat(), mat) < 5e-3, tag + "two_sided_id (from ColumnID) error");

    auto rid = lq.compress(CompressionType::ADAPTIVE(1e-4)).row_id();    // src/row_interp_decomp.rs:176-224
    expect(rel_diff_fro(rid.to_mat(), mat) < 5e-4 * std::sqrt((double)kmin), tag + "row_id error");
    expect(rel_diff_fro(rid.two_sided_id().to_mat(), mat) < 5e-3, tag + "two_sided_id (from RowID) error");

    auto svd = SVD<A>::compute_from(mat);                                 // src/svd.rs:200-290
    expect(rel_diff_fro(svd.to_mat(), mat) < tol, tag + "svd reconstruction");
    auto s = svd.get_s();
    bool sorted = true;
    for (size_t i = 1; i < s.size(); ++i) sorted = sorted && s[i] <= s[i - 1];
    expect(sorted && std::fabs((double)s[0] - 1.0) < 100 * tol, tag + "singular values descending, sigma_0 = 1");
    expect(svd.compress(CompressionType::RANK(20)).rank() == 20, tag + "compress_svd_rank");
    expect(rel_diff_fro(svd.compress(CompressionType::ADAPTIVE(1e-4)).to_mat(), mat) < 5e-4 * std::sqrt((double)kmin),
           tag + "compress_svd_tolerance error");
    expect(rel_diff_fro(svd.to_qr().to_mat(), mat) < tol * 10, tag + "svd.to_qr reconstruction");

    // containers assembled from parts (pub fields of QR / LQ / SVD: src/qr.rs:31-51, src/svd.rs:13-20) behave like the originals
    expect(rel_diff_fro(QR<A>::make(qr.get_q(), qr.get_r(), qr.get_ind()).to_mat(), mat) < tol, tag + "QR{q,r,ind} from parts");
    expect(rel_diff_fro(LQ<A>::make(lq.get_l(), lq.get_q(), lq.get_ind()).to_mat(), mat) < tol, tag + "LQ{l,q,ind} from parts");
    expect(rel_diff_fro(SVD<A>::make(svd.get_u(), svd.get_s(), svd.get_vt()).to_mat(), mat) < tol, tag + "SVD{u,s,vt} from parts");
    expect(rel_diff_fro(rid.dot(x_for_apply(ctx, cols, A(0))), rid.to_mat().matmat(x_for_apply(ctx, cols, A(0)))) < tol * 100, tag + "RowID::dot");
    // Apply: ID * matrix equals to_mat * matrix
    auto x = Matrix<A>::random_gaussian(ctx, cols, 3, 5);
    expect(rel_diff_fro(cid.dot(x), cid.to_mat().matmat(x)) < tol * 100, tag + "ColumnID::dot");
}

static void permutation_known_answers(const Context& ctx) {               // src/permutation.rs:192-239
    std::vector<size_t> perm = {1, 2, 0};
    auto inv = invert_permutation_vector(perm);
    expect(inv == std::vector<size_t>({2, 0, 1}), "invert_permutation_vector [1,2,0] -> [2,0,1]");
    std::vector<double> m = {1, 2, 3, 4, 5, 6, 7, 8, 9};
    auto mat = Matrix<double>::from_host(ctx, m, 3, 3);
    auto col = apply_permutation(mat, perm, MatrixPermutationMode::COL).to_host();      // out[:, i] = in[:, p[i]]
    expect(col == std::vector<double>({2, 3, 1, 5, 6, 4, 8, 9, 7}), "apply_permutation COL");
    auto row = apply_permutation(mat, perm, MatrixPermutationMode::ROW).to_host();      // out[i, :] = in[p[i], :]
    expect(row == std::vector<double>({4, 5, 6, 7, 8, 9, 1, 2, 3}), "apply_permutation ROW");
    auto colinv = apply_permutation(apply_permutation(mat, perm, MatrixPermutationMode::COL), perm, MatrixPermutationMode::COLINV).to_host();
    expect(colinv == m, "COLINV undoes COL");
    auto rowinv = apply_permutation(apply_permutation(mat, perm, MatrixPermutationMode::ROW), perm, MatrixPermutationMode::ROWINV).to_host();
    expect(rowinv == m, "ROWINV undoes ROW");
    bool threw = false;
    // the crate indexes `inverse[elem]` unchecked against duplicates (src/permutation.rs:33-35): only an out-of-range
    // entry panics there, which the ABI reports as RC_INVALID_ARGUMENT
    try { invert_permutation_vector({0, 3, 1}); } catch (const InvalidArgument&) { threw = true; }
    expect(threw, "invert_permutation_vector rejects an out-of-range entry");
    expect(invert_permutation_vector({0, 0, 1}).size() == 3, "invert_permutation_vector accepts duplicates like the crate");
}

int main() {
    Context ctx(0);
    permutation_known_answers(ctx);
    const size_t shapes[2][2] = {{100, 50}, {50, 100}};
    for (auto& sh : shapes) {
        factorization_tests<double>(ctx, "f64", 1e-10, sh[0], sh[1]);
        factorization_tests<c64>(ctx, "c64", 1e-10, sh[0], sh[1]);
        factorization_tests<float>(ctx, "f32", 1e-4, sh[0], sh[1]);
        factorization_tests<c32>(ctx, "c32", 1e-4, sh[0], sh[1]);
    }
    // samplers over the operator traits (src/random_sampling.rs): adaptive range finder then QR from the range estimate
    auto mat = Matrix<double>::random_approximate_low_rank_matrix(ctx, 500, 200, 1.0, 1e-10, 0);
    auto qh = sample_range_adaptive(mat, 1e-5, 5, 1);
    expect(qh.first.ncols() >= 100 && qh.first.ncols() <= 130 && qh.second.back().second < 1e-5, "sample_range_adaptive rank ~115");
    expect(rel_diff_fro(QR<double>::compute_from_range_estimate(qh.first, mat).to_mat(), mat) < 5e-5, "QR::compute_from_range_estimate");
    auto q = sample_range_power_iteration(mat, 40, 10, 2, 3);
    auto svd = SVD<double>::compute_from_range_estimate(q, mat);
    expect(svd.rank() == 40 && rel_diff_fro(svd.to_mat(), mat) < 1e-1, "sample_range_power_iteration + SVD::compute_from_range_estimate");
    expect(max_col_norm(q) < 1.0 + 1e-12 && max_col_norm(q) > 1.0 - 1e-12, "max_col_norm of an orthonormal basis");
    // pipelined upload (rc_matrix_from_host_async / rc_matrix_await): two operators in flight, same bytes as the blocking route
    {
        std::vector<double> h0(300 * 40), h1(300 * 40);
        for (size_t i = 0; i < h0.size(); ++i) { h0[i] = std::sin(0.37 * (double)i); h1[i] = std::cos(0.11 * (double)i); }
        auto a0 = Matrix<double>::from_host_async(ctx, h0.data(), 300, 40, 40);
        auto a1 = Matrix<double>::from_host_async(ctx, h1.data(), 300, 40, 40);
        expect(a0.await_upload().to_host() == h0 && a1.await_upload(true).to_host() == h1, "from_host_async / await_upload");
    }
    std::printf("%d failure(s); %lld kernel launches\n", failures, (long long)ctx.counter("kernel_launches"));
    return failures == 0 ? 0 : 1;
}
