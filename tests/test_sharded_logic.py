"""World-size-2 gloo test (CPU) of the row-sharded algorithm the C++ host code implements
(csrc/host_api.cu: DistTsqr, conj_matmat_impl): local sketch, TSQR with an all-gather of the R
factors, pivoting on the combined R, all-reduce of the A^H Q partials.  Each rank holds a row
block of A; the local numerics are the oracle's numpy/LAPACK calls; the collectives are real
torch.distributed (gloo) calls.  The sharded result must reproduce the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_path as ref
from oracle.inputs import decaying_spectrum_matrix
from oracle.philox import random_gaussian

M, N, K, P = 600, 200, 24, 6


def _allgather(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [o.numpy() for o in outs]


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x).copy())
    dist.all_reduce(t)
    return t.numpy()


def sharded_pivoted_qr(y_local, ncq):
    """DistTsqr + pivot-on-R: local QR, all-gather R_i, QR of the stack (redundant on every rank),
    pivoted QR of the combined R, Q_i = Q0_i * top-chunk_i * Q1."""
    rank = dist.get_rank()
    w = y_local.shape[1]
    q0, r0 = np.linalg.qr(y_local)                      # local TSQR (Householder)
    rs = _allgather(r0)
    qs, r = np.linalg.qr(np.concatenate(rs, axis=0))    # top level on every rank
    q1, rr, ind = ref.pivoted_qr(r)                     # pivoting on the small factor
    top = qs[rank * w:(rank + 1) * w, :].dot(q1[:, :ncq])
    return q0.dot(top), rr, ind


def sharded_rsvd(a_local, omega, k, it_count):
    y0 = a_local.dot(omega)
    res = y0
    for index in range(it_count):
        q, _, _ = sharded_pivoted_qr(y0, y0.shape[1])                 # always y0 (quirk Q1)
        z = _allreduce(np.conj(a_local.T).dot(q))                      # all-reduce of A^H Q partials
        w, _, _ = ref.pivoted_qr(z)                                   # replicated n x l: plain pivoted QR
        ynew = a_local.dot(w)
        if index == it_count - 1:
            res = ynew
    q, _, ind = sharded_pivoted_qr(res, k)
    b = np.conj(_allreduce(np.conj(a_local.T).dot(q)).T)              # b = Q^H A, replicated
    ub, s, vt = ref.compute_svd(b)
    return q.dot(ub), s, vt, ind


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, _ = decaying_spectrum_matrix(M, N, np.float64, seed=5, r0=64, decade_every=8.0)
    omega = random_gaussian((N, K + P), np.float64, seed=42)           # same Philox seed on every rank
    rows = M // world
    a_local = a[rank * rows:(rank + 1) * rows]
    u_local, s, vt, ind = sharded_rsvd(a_local, omega, K, it_count=2)
    us = _allgather(u_local)
    if rank == 0:
        np.savez(out, u=np.concatenate(us, axis=0), s=s, vt=vt, ind=ind)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_row_sharded_rsvd_matches_unsharded_oracle(tmp_path):
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    a, _ = decaying_spectrum_matrix(M, N, np.float64, seed=5, r0=64, decade_every=8.0)
    omega = random_gaussian((N, K + P), np.float64, seed=42)
    q_ref = ref.sample_range_power_iteration(a, K, P, 2, ref.OmegaStream(np.float64, blocks=[omega]))
    svd_ref = ref.SVD.compute_from_range_estimate(q_ref, a)
    assert np.max(np.abs(got["s"] - svd_ref.s) / svd_ref.s) < 1e-10
    u = got["u"]
    assert np.max(np.abs(u.T.dot(u) - np.eye(K))) < 1e-12
    rec = (u * got["s"]).dot(got["vt"])
    assert abs(ref.rel_diff_fro(rec, a) - ref.rel_diff_fro(svd_ref.to_mat(), a)) < 1e-10
    # pivots of the final sketch agree with the unsharded LAPACK path
    y = a.dot(ref.pivoted_qr(np.conj(a.T).dot(ref.pivoted_qr(a.dot(omega))[0]))[0])
    assert np.array_equal(got["ind"], ref.pivoted_qr(y)[2])


# ---- row-sharded column ID + two-sided ID (csrc/host_api.cu: rc_column_id_two_sided_id, sharded branch)
def sharded_two_sided_id(a_local, omega, k):
    """by-rank sampling (sharded pivoted QR of the sketch) -> b = Q^H A by all-reduce, pivoted QR of b replicated
    -> C = Q (q_b R11) local rows, Z replicated -> C all-gathered, pivoted LQ + row ID replicated, every rank
    keeps its own rows of X."""
    rank, world = dist.get_rank(), dist.get_world_size()
    q, _, _ = sharded_pivoted_qr(a_local.dot(omega), k)
    b = np.conj(_allreduce(np.conj(a_local.T).dot(q)).T)               # k x n, replicated
    qb, rb, ind = ref.pivoted_qr(b)
    qr = ref.QR(q.dot(qb), rb, ind).compress(ref.RANK(k))              # q holds the local rows
    cid = qr.column_id()                                               # C local rows, Z replicated
    c_full = np.concatenate(_allgather(cid.c), axis=0)                 # all-gather of C (m x k)
    rid = ref.LQ.compute_from(c_full).row_id()                         # replicated on every rank
    rows = a_local.shape[0]
    x_local = rid.x[rank * rows:(rank + 1) * rows]
    return x_local, rid.r, cid.z, rid.row_ind, cid.col_ind


def _worker_id(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, _ = decaying_spectrum_matrix(M, N, np.float64, seed=6, r0=64, decade_every=8.0)
    omega = random_gaussian((N, K + P), np.float64, seed=43)
    rows = M // world
    x_local, xr, z, row_ind, col_ind = sharded_two_sided_id(a[rank * rows:(rank + 1) * rows], omega, K)
    xs = _allgather(x_local)
    if rank == 0:
        np.savez(out, c=np.concatenate(xs, axis=0), x=xr, r=z, row_ind=row_ind, col_ind=col_ind)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_two_sided_id_matches_unsharded_oracle(tmp_path):
    out = str(tmp_path / "sharded_id.npz")
    mp.spawn(_worker_id, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    a, _ = decaying_spectrum_matrix(M, N, np.float64, seed=6, r0=64, decade_every=8.0)
    omega = random_gaussian((N, K + P), np.float64, seed=43)
    q_ref = ref.sample_range_by_rank(a, K, P, ref.OmegaStream(np.float64, blocks=[omega]))
    ts_ref = ref.QR.compute_from_range_estimate(q_ref, a).compress(ref.RANK(K)).column_id().two_sided_id()
    assert np.array_equal(got["col_ind"][:K], ts_ref.col_ind[:K])
    assert np.array_equal(got["row_ind"][:K], ts_ref.row_ind[:K])
    e = ref.rel_diff_fro(got["c"].dot(got["x"].dot(got["r"])), a)
    e_ref = ref.rel_diff_fro(ts_ref.to_mat(), a)
    assert abs(e - e_ref) <= 1e-8 * e_ref


# ---- row-sharded Cholesky-QR chain (csrc/host_api.cu: cholqr_rounds, cholqr2_acceptable): the Gram matrices are
# all-reduced, so every rank sees the same status words and takes the same route (plain -> shifted -> Householder)
U_F64 = 1.1102230246251565e-16


def _chol_upper(g):
    """(R, breakdown flag, min diag, max diag) of G = R^H R; never raises (a breakdown is a status, like on the device)."""
    try:
        r = np.conj(np.linalg.cholesky(g).T)
    except np.linalg.LinAlgError:
        return np.eye(g.shape[0]), 1.0, 0.0, 0.0
    d = np.abs(np.diag(r))
    return r, 0.0, float(d.min()), float(d.max())


def sharded_cholqr_rounds(y_local, m_glob, shifted):
    """Returns (q_local, R, status) with status = the 16 words the device collects."""
    w = y_local.shape[1]
    h = np.zeros(16)
    r0 = np.eye(w)
    y = y_local
    if shifted:
        s_rel = 100.0 * U_F64 * (np.sqrt(m_glob) + w) * w
        for slot in (8, 12):                                             # the two shifted rounds
            g = _allreduce(np.conj(y.T).dot(y))
            g = g + s_rel * np.max(np.real(np.diag(g))) * np.eye(w)
            r, h[slot], h[slot + 1], h[slot + 2] = _chol_upper(g)
            y = y.dot(np.linalg.inv(r))
            r0 = r.dot(r0)
    g1 = _allreduce(np.conj(y.T).dot(y))
    r1, h[0], h[1], h[2] = _chol_upper(g1)
    q1 = y.dot(np.linalg.inv(r1))
    g2 = _allreduce(np.conj(q1.T).dot(q1))
    h[7] = np.max(np.abs(g2 - np.eye(w)))
    r2, h[4], h[5], h[6] = _chol_upper(g2)
    return q1.dot(np.linalg.inv(r2)), r2.dot(r1).dot(r0), h


def cholqr2_acceptable(h, shifted):
    """csrc/host_api.cu: cholqr2_acceptable (double precision)."""
    ok0 = (not shifted) or (h[8] == 0.0 and h[9] > 0.0 and h[12] == 0.0 and h[13] > 0.0)
    ok2 = h[4] == 0.0 and h[5] > 0.0 and h[7] <= 0.25
    ok1 = h[0] == 0.0 and h[1] > 0.0 and (h[2] / h[1] <= 1.0e6 or (ok2 and h[7] <= 1.0e-3 and h[6] / h[5] <= 2.0))
    return ok0 and ok1 and ok2


def sharded_tall_qr(y_local, m_glob):
    """The chain of cholqr2(): plain, then shifted; the route taken is returned with the factors."""
    for shifted in (False, True):
        q, r, h = sharded_cholqr_rounds(y_local, m_glob, shifted)
        if cholqr2_acceptable(h, shifted):
            return q, r, "shifted" if shifted else "plain"
    return None, None, "householder"


def _worker_cholqr(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, n, l = 2048, 512, 74
    a, _ = decaying_spectrum_matrix(m, n, np.float64, seed=77, r0=128, decade_every=6.0)     # twelve decades over 74 columns
    omega = random_gaussian((n, l), np.float64, seed=42)
    rows = m // world
    a_local = a[rank * rows:(rank + 1) * rows]
    y0 = a_local.dot(omega)                                              # not graded, cond ~1e13: plain route must be rejected
    q0, r0, route0 = sharded_tall_qr(y0, m)
    q1piv, _, _ = ref.pivoted_qr(r0)                                     # pivot on R, Q = Q0 Q1 (local rows)
    z = _allreduce(np.conj(a_local.T).dot(q0.dot(q1piv)))                # graded like the spectrum: plain route, rule (b)
    rows_z = n // world
    qz, rz, route_z = sharded_tall_qr(z[rank * rows_z:(rank + 1) * rows_z], n)
    routes = [None] * world
    dist.all_gather_object(routes, (route0, route_z))
    q0_full, qz_full = np.concatenate(_allgather(q0), axis=0), np.concatenate(_allgather(qz), axis=0)
    if rank == 0:
        np.savez(out, q0=q0_full, r0=r0, qz=qz_full, rz=rz, z=z, routes=np.array([f"{a_}/{b_}" for a_, b_ in routes]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_cholqr_routes_agree_across_ranks(tmp_path):
    out = str(tmp_path / "sharded_cholqr.npz")
    mp.spawn(_worker_cholqr, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert list(got["routes"]) == ["shifted/plain", "shifted/plain"]       # same decision on every rank
    m, n, l = 2048, 512, 74
    a, _ = decaying_spectrum_matrix(m, n, np.float64, seed=77, r0=128, decade_every=6.0)
    omega = random_gaussian((n, l), np.float64, seed=42)
    y0 = a.dot(omega)
    q0, r0 = got["q0"], got["r0"]
    assert np.max(np.abs(q0.T.dot(q0) - np.eye(l))) < 1e-13
    assert np.linalg.norm(q0.dot(r0) - y0) <= 1e-14 * np.linalg.norm(y0)
    # |diag| of the pivoted R against the unsharded LAPACK path, to the roundoff of ||Y0||
    d, d_ref = np.abs(np.diag(ref.pivoted_qr(r0)[1])), np.abs(np.diag(ref.pivoted_qr(y0)[1]))
    assert np.max(np.abs(d - d_ref)) <= 1e-13 * d_ref[0]
    # the graded sketch on the plain route: COLUMNWISE backward stable (every column to its own norm) although diag(R1)
    # spans twelve decades
    qz, rz, z = got["qz"], got["rz"], got["z"]
    assert np.max(np.abs(qz.T.dot(qz) - np.eye(l))) < 1e-13
    col = np.linalg.norm(qz.dot(rz) - z, axis=0) / np.linalg.norm(z, axis=0)
    assert np.max(col) < 1e-12 and np.linalg.norm(z, axis=0).min() < 1e-9 * np.linalg.norm(z, axis=0).max()
