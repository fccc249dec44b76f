"""ORACLE -- test infrastructure, not product code (only tests/ may import it).

numpy float64 restatement of the Sim(3) pose-graph Levenberg-Marquardt step of the reference backend
(SURVEY.md section 8(f) rank 4):

  * residual      vista_slam/pose_graph.py:70-154  PoseGraphOpt.forward:  r_e = Log(T_e * X_i^-1 * X_j)  in R^7
                  with the optimised / fixed node split of :73-98 and the related-edge mask of :150-154
  * LM step       vista_slam/slam.py:108-140  pp.optim.LM(graph, solver=Cholesky, strategy=TrustRegion(radius=1e4),
                  min=1e-6, vectorize=True), weight = diag(conf_e) (7 per edge), StopOnPlateau(steps=20, patience=3,
                  decreasing=1e-4)

The arithmetic lives in PyPose (requirements.txt:10, version unpinned), which is ABSENT from /root/reference and from this
container: **parity unpinned** against the third-party code.  The restatement follows PyPose's published conventions
(LieTensor docs): Sim3 data = [tx ty tz | qx qy qz qw | s] acting as x -> s R x + t; sim3 tangent = [tau | phi | sigma];
Exp / Log are the matrix exponential / logarithm of the 4x4 generator [[sigma I + [phi]x, tau], [0, 0]]; parameters are
updated by LEFT multiplication X <- Exp(delta) X (LieTensor.add_), so the Jacobian is taken w.r.t. a left perturbation.
It is pinned by closed-form known answers instead (tests/test_pose_graph.py): scipy's matrix exponential of the
generator, hand-computed group elements, Log(Exp(xi)) = xi, analytic vs central-difference Jacobians, and convergence of
the LM iteration to the planted solution of a synthetic graph.
"""
import numpy as np
from scipy.linalg import expm


# ---------------------------------------------------------------------------------------------- quaternions
def quat_to_rot(q):
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)  # fp32-rounded inputs are not exactly unit
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    """(x, y, z, w) with w >= 0 (Shepperd's method)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


# ---------------------------------------------------------------------------------------------- Sim(3) group
def sim3_matrix(X):
    """[t | q | s] -> 4x4 [[s R, t], [0, 1]]."""
    M = np.eye(4)
    M[:3, :3] = X[7] * quat_to_rot(X[3:7])
    M[:3, 3] = X[:3]
    return M


def sim3_from_matrix(M):
    sR = M[:3, :3]
    s = np.cbrt(np.linalg.det(sR))
    return np.concatenate([M[:3, 3], rot_to_quat(sR / s), [s]])


def sim3_mul(A, B):
    return sim3_from_matrix(sim3_matrix(A) @ sim3_matrix(B))


def sim3_inv(X):
    return sim3_from_matrix(np.linalg.inv(sim3_matrix(X)))


def sim3_generator(xi):
    G = np.zeros((4, 4))
    G[:3, :3] = xi[6] * np.eye(3) + hat(xi[3:6])
    G[:3, 3] = xi[:3]
    return G


def sim3_exp(xi):
    """Closed form (Eade, "Lie groups for 2D and 3D transformations", Sim(3)): s = e^sigma, R = Exp(phi), t = W tau."""
    tau, phi, sigma = xi[:3], xi[3:6], xi[6]
    th = np.linalg.norm(phi)
    s = np.exp(sigma)
    Phi = hat(phi)
    if th < 1e-8:
        R = np.eye(3) + Phi
    else:
        R = np.eye(3) + np.sin(th) / th * Phi + (1 - np.cos(th)) / th ** 2 * Phi @ Phi
    # W = C I + A Phi + B Phi^2 ; robust evaluation through the integral  W = int_0^1 exp(u (sigma I + Phi)) du
    # (the closed-form coefficients are singular at sigma -> 0 and theta -> 0; the oracle can afford the quadrature-free
    # matrix function instead): phi_1 of the 3x3 generator block
    Gb = sigma * np.eye(3) + Phi
    aug = np.zeros((6, 6))
    aug[:3, :3] = Gb
    aug[:3, 3:] = np.eye(3)
    W = expm(aug)[:3, 3:]
    return np.concatenate([W @ tau, rot_to_quat(R), [s]])


def sim3_log(X):
    """Inverse of sim3_exp (rotation angle in [0, pi))."""
    t, q, s = X[:3], X[3:7], X[7]
    sigma = np.log(s)
    q = q / np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    n = np.linalg.norm(q[:3])
    if n < 1e-12:
        phi = 2.0 * q[:3]
    else:
        phi = 2.0 * np.arctan2(n, q[3]) * q[:3] / n
    Gb = sigma * np.eye(3) + hat(phi)
    aug = np.zeros((6, 6))
    aug[:3, :3] = Gb
    aug[:3, 3:] = np.eye(3)
    W = expm(aug)[:3, 3:]
    return np.concatenate([np.linalg.solve(W, t), phi, [sigma]])


def sim3_adj(X):
    """7x7 adjoint: Exp(Ad_X xi) = X Exp(xi) X^-1, tangent order [tau | phi | sigma]."""
    t, R, s = X[:3], quat_to_rot(X[3:7]), X[7]
    A = np.zeros((7, 7))
    A[:3, :3] = s * R
    A[:3, 3:6] = hat(t) @ R
    A[:3, 6] = -t
    A[3:6, 3:6] = R
    A[6, 6] = 1.0
    return A


def sim3_ad(xi):
    """7x7 ad_xi (Lie bracket [xi, .]) -- the derivative of Ad_{Exp(u xi)} at u = 0."""
    tau, phi, sigma = xi[:3], xi[3:6], xi[6]
    a = np.zeros((7, 7))
    a[:3, :3] = sigma * np.eye(3) + hat(phi)
    a[:3, 3:6] = hat(tau)
    a[:3, 6] = -tau
    a[3:6, 3:6] = hat(phi)
    return a


def sim3_jl(xi):
    """Left Jacobian J_l(xi) = sum_n ad^n / (n+1)! = phi_1(ad_xi), evaluated as a matrix function (no series cut-off)."""
    a = sim3_ad(xi)
    aug = np.zeros((14, 14))
    aug[:7, :7] = a
    aug[:7, 7:] = np.eye(7)
    return expm(aug)[:7, 7:]


def sim3_jl_inv(xi):
    return np.linalg.inv(sim3_jl(xi))


# ---------------------------------------------------------------------------------------------- pose graph
def edge_residual(T_e, X_i, X_j):
    """pose_graph.py:143-149: Log(T_e @ X_i.Inv() @ X_j) in R^7."""
    return sim3_log(sim3_mul(sim3_mul(T_e, sim3_inv(X_i)), X_j))


def edge_jacobians(T_e, X_i, X_j):
    """(r, dr/d delta_i, dr/d delta_j) for left perturbations X <- Exp(delta) X:
    E' = T X_i^-1 Exp(d_j) X_j = Exp(Ad_A d_j) E with A = T X_i^-1, hence dr/dd_j = J_l^-1(r) Ad_A and dr/dd_i = -dr/dd_j."""
    A = sim3_mul(T_e, sim3_inv(X_i))
    r = sim3_log(sim3_mul(A, X_j))
    Jj = sim3_jl_inv(r) @ sim3_adj(A)
    return r, -Jj, Jj


def split_nodes(num_nodes, opt_idx):
    """pose_graph.py:73-98: (opt_map, is_opt) -- local index of every optimised node, -1 for fixed ones."""
    opt_map = -np.ones(num_nodes, dtype=np.int64)
    opt_map[np.asarray(opt_idx, dtype=np.int64)] = np.arange(len(opt_idx))
    return opt_map


def related_edges(edges, opt_idx):
    """pose_graph.py:150-154: edges with at least one optimised endpoint."""
    m = np.isin(edges, np.asarray(opt_idx))
    return m[:, 0] | m[:, 1]


def build_normal_equations(nodes, edges, meas, weights, opt_idx):
    """A = J^T W J (7 n_opt square), g = J^T W r, loss = r^T W r over the related edges; fixed nodes contribute no columns
    (pose_graph.py:100-149: three edge classes, all handled by the opt_map lookup)."""
    opt_map = split_nodes(len(nodes), opt_idx)
    n = len(opt_idx)
    A = np.zeros((7 * n, 7 * n))
    g = np.zeros(7 * n)
    loss = 0.0
    for e in np.nonzero(related_edges(edges, opt_idx))[0]:
        i, j = int(edges[e, 0]), int(edges[e, 1])
        r, Ji, Jj = edge_jacobians(meas[e], nodes[i], nodes[j])
        W = np.diag(weights[e])
        loss += float(r @ W @ r)
        blocks = [(opt_map[i], Ji), (opt_map[j], Jj)]
        for a, Ja in blocks:
            if a < 0:
                continue
            g[7 * a:7 * a + 7] += Ja.T @ W @ r
            for b, Jb in blocks:
                if b < 0:
                    continue
                A[7 * a:7 * a + 7, 7 * b:7 * b + 7] += Ja.T @ W @ Jb
    return A, g, loss


def graph_loss(nodes, edges, meas, weights, opt_idx):
    loss = 0.0
    for e in np.nonzero(related_edges(edges, opt_idx))[0]:
        r = edge_residual(meas[e], nodes[int(edges[e, 0])], nodes[int(edges[e, 1])])
        loss += float(r @ (weights[e] * r))
    return loss


def apply_update(nodes, opt_idx, delta):
    out = nodes.copy()
    for a, v in enumerate(opt_idx):
        out[v] = sim3_mul(sim3_exp(delta[7 * a:7 * a + 7]), nodes[v])
    return out


def lm_step(nodes, edges, meas, weights, opt_idx, damping, dmin=1e-6, dmax=1e32):
    """One damped Gauss-Newton solve of pp.optim.LM.step: A = J^T W J with its diagonal clamped to [min, max], then
    A.diagonal() += damping * A.diagonal(), Cholesky solve of A d = -J^T W r, left-multiplicative update.
    Returns (new_nodes, loss_before, loss_after, delta)."""
    A, g, loss0 = build_normal_equations(nodes, edges, meas, weights, opt_idx)
    d = np.clip(np.diag(A).copy(), dmin, dmax)
    A[np.diag_indices_from(A)] = d * (1.0 + damping)
    L = np.linalg.cholesky(A)
    delta = -np.linalg.solve(L.T, np.linalg.solve(L, g))
    new_nodes = apply_update(nodes, opt_idx, delta)
    return new_nodes, loss0, graph_loss(new_nodes, edges, meas, weights, opt_idx), delta


def optimize(nodes, edges, meas, weights, opt_idx, steps=20, patience=3, decreasing=1e-4, radius=1e4):
    """slam.py:121-134: LM with a trust-region damping (damping = 1 / radius, radius adapted by the step quality as in
    pp.optim.strategy.TrustRegion: high = 0.5, low = 1e-3, up = 2, down = 0.5) and StopOnPlateau(steps, patience,
    decreasing).  A rejected step (loss went up) is undone and retried with a smaller radius."""
    losses = []
    stall = 0
    for _ in range(steps):
        new_nodes, l0, l1, delta = lm_step(nodes, edges, meas, weights, opt_idx, 1.0 / radius)
        if l1 <= l0:
            nodes = new_nodes
            radius = min(radius * 2.0, 1e32)
        else:
            radius = max(radius * 0.5, 1e-32)
            l1 = l0
        losses.append(l1)
        stall = stall + 1 if (l0 - l1) < decreasing else 0
        if stall >= patience:
            break
    return nodes, losses
