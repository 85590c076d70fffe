"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel once on ragged small sizes."""
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
be.close()
print("sanitize pass done")
