"""Sim(3) pose-graph LM step (vista_slam/pose_graph.py:70-154, slam.py:108-140).

CPU part: pins the numpy oracle by closed-form known answers (PyPose, where the reference's arithmetic lives, is absent:
parity unpinned against the third-party code -- see oracle/pose_graph_oracle.py).
GPU part (-m gpu): the CUDA residual / Jacobian / normal-equation / Cholesky kernels against that oracle.
"""
import numpy as np
import pytest
from scipy.linalg import expm

from oracle import pose_graph_oracle as pg


def rand_sim3(rng, scale=1.0):
    xi = np.concatenate([rng.normal(size=3) * scale, rng.normal(size=3) * 0.7 * scale, rng.normal(size=1) * 0.3 * scale])
    return pg.sim3_exp(xi), xi


def test_exp_is_the_matrix_exponential_of_the_generator():
    rng = np.random.default_rng(0)
    for _ in range(20):
        X, xi = rand_sim3(rng)
        assert np.allclose(pg.sim3_matrix(X), expm(pg.sim3_generator(xi)), atol=1e-12)
    # degenerate corners of the closed form: no rotation, no scale change, neither
    for xi in ([1, 2, 3, 0, 0, 0, 0.4], [1, 2, 3, 0.3, -0.2, 0.5, 0.0], [1, 2, 3, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1e-10, 1e-10]):
        xi = np.array(xi, dtype=float)
        assert np.allclose(pg.sim3_matrix(pg.sim3_exp(xi)), expm(pg.sim3_generator(xi)), atol=1e-12)


def test_hand_computed_group_elements():
    # rotation by 90 degrees about z, nothing else
    X = pg.sim3_exp(np.array([0, 0, 0, 0, 0, np.pi / 2, 0]))
    assert np.allclose(X, [0, 0, 0, 0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4), 1])
    # pure scale + translation: x' = sigma x + tau integrates to s = e^sigma, t = (e^sigma - 1) / sigma * tau
    sig = 0.7
    X = pg.sim3_exp(np.array([1.0, -2.0, 0.5, 0, 0, 0, sig]))
    assert np.allclose(X[7], np.exp(sig)) and np.allclose(X[:3], (np.exp(sig) - 1) / sig * np.array([1.0, -2.0, 0.5]))
    # action x -> s R x + t, composition and inverse
    A = np.array([1.0, 2.0, 3.0, 0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4), 2.0])   # 90 deg about z, scale 2
    x = np.array([1.0, 0.0, 0.0])
    assert np.allclose((pg.sim3_matrix(A) @ np.append(x, 1))[:3], [1.0, 4.0, 3.0])  # 2 * (0,1,0) + (1,2,3)
    AA = pg.sim3_mul(A, A)
    assert np.allclose(AA[7], 4.0) and np.allclose((pg.sim3_matrix(AA) @ np.append(x, 1))[:3], [-7.0, 4.0, 9.0])  # 2 R (1,4,3) + (1,2,3) = (-8,2,6) + (1,2,3)
    I = pg.sim3_mul(A, pg.sim3_inv(A))
    assert np.allclose(I, [0, 0, 0, 0, 0, 0, 1, 1], atol=1e-12)


def test_log_inverts_exp_and_adjoint_identity():
    rng = np.random.default_rng(1)
    for _ in range(30):
        X, xi = rand_sim3(rng)
        assert np.allclose(pg.sim3_log(X), xi, atol=1e-10)
        Y, eta = rand_sim3(rng)
        # X Exp(eta) X^-1 = Exp(Ad_X eta)
        lhs = pg.sim3_matrix(pg.sim3_mul(pg.sim3_mul(X, Y), pg.sim3_inv(X)))
        rhs = expm(pg.sim3_generator(pg.sim3_adj(X) @ eta))
        assert np.allclose(lhs, rhs, atol=1e-10)
    assert np.allclose(pg.sim3_log(np.array([0, 0, 0, 0, 0, 0, 1, 1.0])), 0)


def test_left_jacobian_against_its_series_and_finite_differences():
    rng = np.random.default_rng(2)
    xi = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 0.2, [0.1]])
    a = pg.sim3_ad(xi)
    series, term = np.zeros((7, 7)), np.eye(7)
    for n in range(30):
        series += term
        term = term @ a / (n + 2)
    assert np.allclose(pg.sim3_jl(xi), series, atol=1e-12)
    # Log(Exp(eps) Exp(xi)) = xi + J_l^-1(xi) eps + O(eps^2)
    X = pg.sim3_exp(xi)
    J = np.zeros((7, 7))
    h = 1e-6
    for k in range(7):
        e = np.zeros(7)
        e[k] = h
        J[:, k] = (pg.sim3_log(pg.sim3_mul(pg.sim3_exp(e), X)) - pg.sim3_log(pg.sim3_mul(pg.sim3_exp(-e), X))) / (2 * h)
    assert np.allclose(J, pg.sim3_jl_inv(xi), atol=1e-7)


def test_edge_jacobians_against_central_differences():
    rng = np.random.default_rng(3)
    for _ in range(5):
        (T, _), (Xi, _), (Xj, _) = rand_sim3(rng, 0.5), rand_sim3(rng), rand_sim3(rng)
        r, Ji, Jj = pg.edge_jacobians(T, Xi, Xj)
        assert np.allclose(r, pg.edge_residual(T, Xi, Xj))
        h = 1e-6
        for which, J in ((0, Ji), (1, Jj)):
            num = np.zeros((7, 7))
            for k in range(7):
                e = np.zeros(7)
                e[k] = h
                P, M = pg.sim3_exp(e), pg.sim3_exp(-e)
                if which == 0:
                    num[:, k] = (pg.edge_residual(T, pg.sim3_mul(P, Xi), Xj) - pg.edge_residual(T, pg.sim3_mul(M, Xi), Xj)) / (2 * h)
                else:
                    num[:, k] = (pg.edge_residual(T, Xi, pg.sim3_mul(P, Xj)) - pg.edge_residual(T, Xi, pg.sim3_mul(M, Xj))) / (2 * h)
            assert np.allclose(num, J, atol=2e-6), (which, np.abs(num - J).max())


def make_graph(rng, n_nodes=12, n_loops=4, noise=0.05, meas_noise=0.0):
    """A planted Sim(3) trajectory, odometry + loop edges T_e = X_j^-1 * X_i (slam.py:215-217: pose_i = pose_j @ sim3_ij, edge
    (i, j, sim3_ij)), so that Log(T_e X_i^-1 X_j) = 0 at the solution, and a perturbed initial guess."""
    gt = [np.array([0, 0, 0, 0, 0, 0, 1, 1.0])]
    for _ in range(n_nodes - 1):
        step, _ = rand_sim3(rng, 0.3)
        gt.append(pg.sim3_mul(gt[-1], step))
    gt = np.array(gt)
    edges = [(i + 1, i) for i in range(n_nodes - 1)] + [(i + 2, i) for i in range(n_nodes - 2)]
    for _ in range(n_loops):
        i, j = sorted(rng.choice(n_nodes, size=2, replace=False))
        edges.append((int(j), int(i)))
    edges = np.array(edges, dtype=np.int64)
    meas = []
    for i, j in edges:
        T = pg.sim3_mul(pg.sim3_inv(gt[j]), gt[i])
        if meas_noise > 0:
            T = pg.sim3_mul(pg.sim3_exp(rng.normal(size=7) * meas_noise), T)
        meas.append(T)
    meas = np.array(meas)
    weights = 0.5 + rng.random((len(edges), 7))
    init = gt.copy()
    for v in range(1, n_nodes):
        init[v] = pg.sim3_mul(pg.sim3_exp(rng.normal(size=7) * noise), gt[v])
    return gt, init, edges, meas, weights


def test_lm_recovers_a_planted_graph_with_fixed_nodes():
    rng = np.random.default_rng(4)
    gt, init, edges, meas, weights = make_graph(rng)
    opt_idx = list(range(1, len(gt)))          # node 0 fixed (gauge), as the reference's window leaves old nodes fixed
    assert pg.related_edges(edges, opt_idx).all()
    nodes, losses = pg.optimize(init, edges, meas, weights, opt_idx, steps=20)
    assert losses[0] > losses[-1] and losses[-1] < 1e-12
    for v in range(len(gt)):
        assert np.allclose(pg.sim3_matrix(nodes[v]), pg.sim3_matrix(gt[v]), atol=1e-5)
    # a window: only the last 4 nodes are optimised, only the edges touching them count, the rest never move
    opt_idx = list(range(len(gt) - 4, len(gt)))
    rel = pg.related_edges(edges, opt_idx)
    assert 0 < rel.sum() < len(edges)
    nodes2, losses2 = pg.optimize(init, edges, meas, weights, opt_idx, steps=20)
    assert np.array_equal(nodes2[:len(gt) - 4], init[:len(gt) - 4])
    assert losses2[-1] < losses2[0]


def test_lm_step_is_a_damped_gauss_newton_solve():
    rng = np.random.default_rng(5)
    gt, init, edges, meas, weights = make_graph(rng, n_nodes=6, n_loops=1, noise=0.02, meas_noise=0.01)
    opt_idx = [1, 2, 3, 4, 5]
    A, g, loss = pg.build_normal_equations(init, edges, meas, weights, opt_idx)
    assert np.allclose(A, A.T) and np.all(np.linalg.eigvalsh(A) > 0) and loss > 0
    _, l0, l1, delta = pg.lm_step(init, edges, meas, weights, opt_idx, 1e-4)
    assert abs(l0 - loss) < 1e-12 and l1 < l0
    Ad = A.copy()
    Ad[np.diag_indices_from(Ad)] *= 1.0 + 1e-4
    assert np.allclose(Ad @ delta, -g, atol=1e-9)


# ------------------------------------------------------------------------------------------------ GPU
def _to_torch(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


@pytest.mark.gpu
def test_gpu_lm_step_matches_the_oracle():
    """one LM step on the device (fp64 inside, fp32 node tensors) vs the numpy oracle on the same fp32-rounded inputs:
    losses, update norm and the updated nodes; windowed problem (fixed + optimised nodes, unrelated edges present)."""
    import torch
    from vista_slam_b200.pose_graph import PoseGraphOpt
    rng = np.random.default_rng(6)
    gt, init, edges, meas, weights = make_graph(rng, n_nodes=40, n_loops=10, noise=0.05, meas_noise=0.02)
    init32, meas32, w32 = init.astype(np.float32), meas.astype(np.float32), weights.astype(np.float32)
    for opt_idx in (list(range(1, 40)), list(range(25, 40))):
        g = PoseGraphOpt(_to_torch(init32), torch.tensor(opt_idx))
        cand, l0, l1, dn, ok = g.lm_step(_to_torch(edges), _to_torch(meas32), _to_torch(w32), 1e-4)
        ref_nodes, r0, r1, delta = pg.lm_step(init32.astype(np.float64), edges, meas32.astype(np.float64), w32.astype(np.float64),
                                              opt_idx, 1e-4)
        assert ok
        assert abs(l0 - r0) < 1e-8 * max(1.0, r0) and abs(l1 - r1) < 1e-6 * max(1.0, r0), (l0, r0, l1, r1)
        assert abs(dn - np.linalg.norm(delta)) < 1e-8 * max(1.0, np.linalg.norm(delta))
        got = cand.cpu().numpy().astype(np.float64)
        for v in range(40):
            assert np.allclose(pg.sim3_matrix(got[v]), pg.sim3_matrix(ref_nodes[v]), atol=2e-6), v   # fp32 output rounding
        rel = g.get_related_edge_idxs(_to_torch(edges)).cpu().numpy()
        assert np.array_equal(rel, pg.related_edges(edges, opt_idx))


@pytest.mark.gpu
def test_gpu_optimisation_recovers_a_planted_graph_and_is_deterministic():
    import torch
    from vista_slam_b200.pose_graph import PoseGraphOpt
    rng = np.random.default_rng(7)
    gt, init, edges, meas, weights = make_graph(rng, n_nodes=60, n_loops=15, noise=0.05)
    opt_idx = torch.arange(1, 60)
    args = (_to_torch(edges), _to_torch(meas, torch.float32), _to_torch(weights, torch.float32))
    g = PoseGraphOpt(_to_torch(init, torch.float32), opt_idx)
    losses = g.optimize(*args)
    assert losses[-1] < 1e-9 and losses[0] > losses[-1]
    got = g.get_nodes().cpu().numpy().astype(np.float64)
    for v in range(60):
        assert np.allclose(pg.sim3_matrix(got[v]), pg.sim3_matrix(gt[v]), atol=2e-4), v
    g2 = PoseGraphOpt(_to_torch(init, torch.float32), opt_idx)
    assert g2.optimize(*args) == losses and torch.equal(g2.get_nodes(), g.get_nodes())   # no atomics anywhere
    # same iteration in the oracle: same loss trajectory
    _, ref_losses = pg.optimize(init.astype(np.float32).astype(np.float64), edges, meas.astype(np.float32).astype(np.float64),
                                weights.astype(np.float32).astype(np.float64), list(range(1, 60)))
    assert len(ref_losses) == len(losses)
    assert abs(ref_losses[0] - losses[0]) < 1e-5 * max(ref_losses[0], 1e-12) + 1e-9
