"""torchrun worker: row-sharded pipelines on N GPUs (one process per GPU) checked against the
unsharded oracle.  Launched by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_path as ref  # noqa: E402
from oracle.inputs import decaying_spectrum_matrix, helmholtz_kernel_matrix  # noqa: E402
from oracle.philox import random_gaussian  # noqa: E402
from rusty_compression_b200 import api  # noqa: E402


def gather_rows(local):
    t = torch.from_numpy(np.ascontiguousarray(local)).cuda()
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy() for o in outs], axis=0)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(device=local)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)

    for dtype, tol in ((np.float64, 1e-10), (np.complex128, 1e-10), (np.float32, 1e-4)):
        m, n, k, p = 4096, 768, 40, 8
        if np.dtype(dtype).kind == "c":
            a = helmholtz_kernel_matrix(m, n, dtype)
        else:
            a, _ = decaying_spectrum_matrix(m, n, dtype, seed=3, r0=128, decade_every=12.0)
        omega = random_gaussian((n, k + p), dtype, seed=42)
        rows = m // world
        op = api.DeviceMatrix.from_numpy(a[rank * rows:(rank + 1) * rows], ctx=ctx).set_shard(m, rank * rows)
        # --- rSVD (power iteration + SVD from range)
        q = api.sample_range_power_iteration(op, k, p, 2, omega=omega, ctx=ctx, device=True)
        svd = api.SVD.compute_from_range_estimate(q, op)
        u = gather_rows(svd.u)
        s, vt = svd.s_f64(), svd.vt
        # --- ID pipeline (by-rank sampling + QR from range + column ID)
        q2 = api.sample_range_by_rank(op, k, p, omega=omega, ctx=ctx, device=True)
        qr = api.QR.compute_from_range_estimate(q2, op).compress(api.RANK(k))
        cid = qr.column_id()
        c_full = gather_rows(cid.c)
        z, col_ind = cid.z, cid.col_ind
        ts = cid.two_sided_id()                      # C is row-sharded: all-gathered, pivoted LQ replicated
        tsc_full = gather_rows(ts.c)
        ts_x, ts_r, ts_row_ind = ts.x, ts.r, ts.row_ind
        # --- adaptive sampler with the Philox stream
        qa, hist = api.sample_range_adaptive(op, 1e-4, 16, seed=7, ctx=ctx, device=True)
        qa_full = gather_rows(qa.to_numpy())
        if rank == 0:
            q_ref = ref.sample_range_power_iteration(a, k, p, 2, ref.OmegaStream(dtype, blocks=[omega]))
            svd_ref = ref.SVD.compute_from_range_estimate(q_ref, a)
            err_s = np.max(np.abs(s - svd_ref.s.astype(np.float64)) / svd_ref.s)
            rec, rec_ref = ref.rel_diff_fro((u * s.astype(u.real.dtype)).dot(vt), a), ref.rel_diff_fro(svd_ref.to_mat(), a)
            assert err_s < tol, (dtype, err_s)
            assert abs(rec - rec_ref) <= tol * rec_ref + (1e-6 if tol > 1e-6 else 0), (dtype, rec, rec_ref)
            q2_ref = ref.sample_range_by_rank(a, k, p, ref.OmegaStream(dtype, blocks=[omega]))
            cid_ref = ref.QR.compute_from_range_estimate(q2_ref, a).compress(ref.RANK(k)).column_id()
            e, e_ref = ref.rel_diff_fro(c_full.dot(z), a), ref.rel_diff_fro(cid_ref.to_mat(), a)
            if np.array_equal(col_ind[:k], cid_ref.col_ind[:k]):
                assert abs(e - e_ref) <= tol * e_ref, (dtype, e, e_ref)
                ts_ref = cid_ref.two_sided_id()
                assert tsc_full.shape == (m, k) and len(ts_row_ind) == m
                e2, e2_ref = ref.rel_diff_fro(tsc_full.dot(ts_x.dot(ts_r)), a), ref.rel_diff_fro(ts_ref.to_mat(), a)
                if np.array_equal(ts_row_ind[:k], ts_ref.row_ind[:k]):
                    assert abs(e2 - e2_ref) <= 10 * tol * e2_ref, (dtype, e2, e2_ref)
                else:
                    assert tol > 1e-6, "f64 row skeleton must match the unsharded LAPACK path"
            else:
                assert tol > 1e-6, "f64 skeleton indices must match the unsharded LAPACK path"
            qa_ref, hist_ref = ref.sample_range_adaptive(a, 1e-4, 16, ref.OmegaStream(dtype, seed=7))
            assert [r for r, _ in hist] == [r for r, _ in hist_ref], (hist, hist_ref)
            ra, ra_ref = ref.range_residual(a, qa_full), ref.range_residual(a, qa_ref)
            assert abs(ra - ra_ref) <= max(tol, 1e-8) * ra_ref + (5e-7 if tol > 1e-6 else 0.0), (dtype, ra, ra_ref)
            print(f"[multi-gpu x{world}] {np.dtype(dtype).name}: sv err {err_s:.2e}, rec {rec:.3e} (ref {rec_ref:.3e}), "
                  f"id err {e:.3e} (ref {e_ref:.3e}), adaptive rank {hist[-1][0]}", flush=True)
    # --- steep spectrum, row-sharded (f64, sigma_j = 10^(-j/6): the 74-column sketch spans twelve decades): the plain
    #     Cholesky-QR2 of Y0 = A Omega is rejected at the early check on every rank alike (the status words come from the
    #     all-reduced Gram matrices) and redone on the shifted Cholesky-QR; the graded sketches of the power iteration pass
    #     by their scaled backward error.  Same tolerances as tests/test_gpu_parity.py::test_rsvd_parity_steep_spectrum.
    m, n, k, p = 4096, 1024, 64, 10
    a, _ = decaying_spectrum_matrix(m, n, np.float64, seed=77, r0=128, decade_every=6.0)
    omega = random_gaussian((n, k + p), np.float64, seed=42)
    rows = m // world
    op = api.DeviceMatrix.from_numpy(a[rank * rows:(rank + 1) * rows], ctx=ctx).set_shard(m, rank * rows)
    ctx.reset_counters()
    q = api.sample_range_power_iteration(op, k, p, 2, omega=omega, ctx=ctx, device=True)
    shifted, rejected = ctx.counter("cholqr_shifted"), ctx.counter("cholqr_fallbacks")
    svd = api.SVD.compute_from_range_estimate(q, op)
    q_full, s = gather_rows(q.to_numpy()), svd.s_f64()
    if rank == 0:
        q_ref = ref.sample_range_power_iteration(a, k, p, 2, ref.OmegaStream(np.float64, blocks=[omega]))
        s_ref = ref.SVD.compute_from_range_estimate(q_ref, a).s
        res, res_ref = ref.range_residual(a, q_full), ref.range_residual(a, q_ref)
        lead = s_ref >= 1e-5 * s_ref[0]
        err_lead, err_abs = np.max(np.abs(s - s_ref)[lead] / s_ref[lead]), np.max(np.abs(s - s_ref)) / s_ref[0]
        orth = np.max(np.abs(q_full.T.dot(q_full) - np.eye(k)))
        print(f"[multi-gpu x{world}] steep spectrum f64 (sharded): shifted Cholesky-QR used {shifted} times after {rejected} rejections, "
              f"|Q^T Q - I| {orth:.2e}, residual {res:.6e} (oracle {res_ref:.6e}), leading sv err {err_lead:.2e}, all sv err / s0 {err_abs:.2e}",
              flush=True)
        assert shifted >= 1 and rejected >= 1
        assert orth < 1e-12 and abs(res - res_ref) <= 1e-3 * res_ref and err_lead <= 1e-10 and err_abs <= 1e-14
    # --- a sketch with more columns than the operator has rank, row-sharded (f64: 300 columns of a rank-100 operator, wider
    #     than one Cholesky / TSQR panel): the later panels have no direction of their own, every rank must take the same
    #     Householder fallback with Gaussian completion, and the gathered range basis must be orthonormal and span range(A)
    rng = np.random.default_rng(17)
    m, n, rk = 8192, 1024, 100
    a = rng.standard_normal((m, rk)).dot(rng.standard_normal((rk, n)))
    rows = m // world
    op = api.DeviceMatrix.from_numpy(a[rank * rows:(rank + 1) * rows], ctx=ctx).set_shard(m, rank * rows)
    q = api.sample_range_by_rank(op, 290, 10, seed=11, ctx=ctx, device=True)
    q_full = gather_rows(q.to_numpy())
    if rank == 0:
        orth = np.max(np.abs(q_full.T.dot(q_full) - np.eye(290)))
        lead = q_full[:, :128]
        res = np.linalg.norm(a - lead.dot(lead.T.dot(a))) / np.linalg.norm(a)
        print(f"[multi-gpu x{world}] rank-100 operator sampled with 300 columns (f64, sharded): |Q^T Q - I| {orth:.2e}, "
              f"residual of the leading 128 columns {res:.2e}", flush=True)
        assert orth < 1e-11 and res < 1e-10
    # --- config-4 shape: row-sharded tall-skinny range finder, f32, rank 256 (+10), n = 8192 (the sketch, l = 266, is wider
    #     than one TSQR panel: panel path with all-reduced projections).  m = 2^19 by default (16 GiB of A, the size SURVEY
    #     8(d) prescribes for the CPU side of config 4); RC_TEST_CONFIG4_ROWS overrides it.
    m, n, k, p = int(os.environ.get("RC_TEST_CONFIG4_ROWS", 1 << 19)), 8192, 256, 10
    rows = m // world
    op = api.tall_shard_matrix(rank * rows, rows, n, np.float32, seed=9, m_total=m, r0=512, decade_every=64.0, ctx=ctx)
    op.set_shard(m, rank * rows)
    q = api.sample_range_by_rank(op, k, p, seed=42, ctx=ctx, device=True)      # Omega regenerated from the shared Philox seed
    q_full = gather_rows(q.to_numpy())
    a_full = gather_rows(op.to_numpy())
    if rank == 0:
        omega = random_gaussian((n, k + p), np.float32, seed=42)
        q_ref = ref.sample_range_by_rank(a_full, k, p, ref.OmegaStream(np.float32, blocks=[omega]))
        # the reference path once more with the sketch formed in double and rounded once (an equally valid f32
        # evaluation of Y = A Omega): how far the reference's OWN result moves under f32 roundoff of the sketch --
        # the deep pivots of the 266-column sketch (trailing norms down to 7e-5 of the first) decide which 10 of the
        # 266 directions are dropped, and they sit at eps_f32 * 1.4e4 relative.  This is the roundoff floor of the
        # comparison; it is printed and the device is held to max(1e-4, 3 x floor).
        y64 = np.zeros((m, k + p), dtype=np.float32)
        for r0 in range(0, m, 1 << 16):
            y64[r0:r0 + (1 << 16)] = a_full[r0:r0 + (1 << 16)].astype(np.float64).dot(omega.astype(np.float64)).astype(np.float32)
        q_alt = ref.QR.compute_from(y64).compress(ref.RANK(k)).q
        cols = np.arange(0, n, 8)                     # same checker, same column sample on every side (f64)
        sub = np.ascontiguousarray(a_full[:, cols])
        r, r_ref, r_alt = (ref.range_residual(sub, x) for x in (q_full, q_ref, q_alt))
        floor = abs(r_alt - r_ref) / r_ref
        orth = np.max(np.abs(q_full.T.astype(np.float64).dot(q_full.astype(np.float64)) - np.eye(k)))
        print(f"[multi-gpu x{world}] config-4 shape f32 {m}x{n}, rank {k}: |Q^T Q - I| {orth:.2e}, residual {r:.6e} "
              f"(oracle {r_ref:.6e}, oracle with the sketch rounded once from double {r_alt:.6e}: roundoff floor {floor:.2e}); "
              f"deviation {abs(r - r_ref) / r_ref:.2e}", flush=True)
        assert orth < 5e-5
        assert abs(r - r_ref) <= max(1e-4, 3.0 * floor) * r_ref, (r, r_ref, floor)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
