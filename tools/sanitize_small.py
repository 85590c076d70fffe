"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once on ragged small sizes.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
(compute-sanitizer is closed on the shared B200 pool of round 2; run plain, the script still touches every kernel on ragged
sizes and prints values that can be compared with the oracle.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from profit_b200 import GPBackend
be = GPBackend(0)
rng = np.random.default_rng(0)
for (N, D, n_ell, kind) in ((37, 3, 3, 0), (300, 5, 1, 1), (260, 10, 10, 1)):
    X = rng.random((N, D)); y = rng.random((N, 2 if N == 37 else 1))
    theta = np.concatenate([np.linspace(0.6, 1.2, D) if n_ell == D else [0.8], [1.5, 0.3]])
    be.set_train(X, y)
    print(N, be.nll(kind, theta), be.nll(kind, theta, want_grad=True)[1][:2])
    be.factorize(kind, theta)
    m, v = be.predict(rng.random((201, D)))
    print("  predict", m.shape, float(v.min()))
    K, dK = be.kernel_matrix(kind, X[:33], X[:70], theta[:-2], 1.5, 0.3, eval_gradient=True)
    print("  K", K.shape, dK.shape)
A = rng.standard_normal((150, 150)); A = A @ A.T / 150 + np.eye(150)
L = be.cholesky(A); Ki = be.invert_cholesky(L); x = be.solve_cholesky(L, rng.random((150, 2)))
print("dense twins", np.abs(Ki @ A - np.eye(150)).max(), be.argmax(rng.random(1000)))
# the truncated eigen-decomposition kernels: Lanczos + QL on a kernel matrix and on a rank-deficient one (restarts), the
# NLL / gradient of the eig exit, Q diag(1/w) Q^T
Z = rng.random((90, 2))
Kz = np.exp(-0.5 * ((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1) / 0.3 ** 2) + 1e-4 * np.eye(90)
w, Q = be.eigsh_matrix(Kz, 7, want_vectors=True)
w2, _ = be.eigsh_matrix(Kz, 21, reuse=True)
print("eigsh", w[-1], w2.shape, be.last_eig_steps, np.abs(be.eig_inverse(np.maximum(w2, 1e-3), 90)).max() > 0)
Xs = np.repeat(rng.random((9, 2)), 3, axis=0); ys = np.sin(Xs[:, 0])
be.set_train(Xs, ys)
ws = be.eigsh(0, np.array([0.5, 1.0, 1e-12]), 12)
print("eig exit", ws[:2], be.nll_eig(24, want_grad=True, n_params=3))
# acquisition and factor extension
X = rng.random((140, 3)); y = rng.random((140, 1)); theta = np.array([0.7, 0.9, 1.1, 1.5, 0.3])
be.set_train(X, y); be.factorize(1, theta)
Xc = rng.random((333, 3)); m, v = be.predict(Xc)
from profit_b200 import backend as B
print("acquire", be.acquire(B.ACQ_EI2, v, term=m, params=np.array([0.01, 0.0, float(y.max())])))
be.append_train(Xc[:2], m[:2]); print("append", be.predict(Xc[:5])[1])
print("marginal variance", be.marginal_variance(Xc[:9], np.eye(5) * 1e-3, v[:9])[:2])
be.close()
print("sanitize pass done")
