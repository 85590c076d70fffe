"""The PyTorch-extension face of the C ABI (profit_b200/csrc/torch_ext.cpp -> torch.ops.profit_b200, north-star item 1):
registration and argument checking without a GPU; parity with the oracle, device-resident outputs and stream ordering
on the GPU."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err


@pytest.fixture(scope="module")
def ops(built_lib):
    from profit_b200 import torch_ops

    return torch_ops.load()


def test_operators_are_registered_with_their_schemas(ops):
    expected = {"create": "profit_b200::create(int device) -> int",
                "set_train": "profit_b200::set_train(int handle, Tensor X, Tensor y) -> ()",
                "nll": "profit_b200::nll(int handle, int kind, Tensor theta, bool want_grad) -> (Tensor, Tensor)",
                "factorize": "profit_b200::factorize(int handle, int kind, Tensor theta) -> ()",
                "predict": "profit_b200::predict(int handle, Tensor Xpred, int add_data_variance, int n_out) -> (Tensor, Tensor)",
                "append_train": "profit_b200::append_train(int handle, Tensor X, Tensor y) -> ()",
                "launch_count": "profit_b200::launch_count(int handle) -> int",
                "destroy": "profit_b200::destroy(int handle) -> ()"}
    for name, schema in expected.items():
        assert str(getattr(ops, name).default._schema) == schema
    assert "kernel_matrix(int handle, int kind, Tensor X, Tensor Y, Tensor length_scale, float sigma_f" in str(
        ops.kernel_matrix.default._schema)


def test_argument_errors_and_no_cpu_fallback(ops):
    with pytest.raises(RuntimeError, match="null handle"):
        ops.launch_count(0)
    with pytest.raises(RuntimeError, match="null handle"):
        ops.nll(0, 0, torch.ones(3, dtype=torch.float64), True)
    # empty kernel matrices return empty tensors like the reference (no device touched)
    K, dK = ops.kernel_matrix(1, 0, torch.zeros(0, 2, dtype=torch.float64), torch.zeros(3, 2, dtype=torch.float64),
                              torch.ones(1, dtype=torch.float64), 1.0, 0.0, True)
    assert tuple(K.shape) == (0, 3) and tuple(dK.shape) == (0, 3, 3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ops.create(0)


@pytest.mark.gpu
def test_torch_ops_parity_and_stream_ordering(ops):
    """NLL + gradient, predict and the dense kernel matrix through torch.ops.profit_b200 against the oracle (north-star
    tolerances), with device tensors produced on a NON-default torch stream and no manual synchronisation."""
    from profit_b200.torch_ops import TorchGP

    N, D, M = 700, 4, 5000
    X, y = oracle.synthetic_problem(N, D)
    theta = np.concatenate([np.linspace(0.6, 1.2, D), [1.5, 0.3]])
    gp = TorchGP(0, "Matern52")
    gp.set_train(torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory())
    nll, grad = gp.nll(torch.from_numpy(theta))
    ref_nll, ref_grad = oracle.nll_cholesky(theta, X, y, oracle.matern52, eval_gradient=True)
    assert nll.device.type == "cpu" and abs(nll.item() - ref_nll) <= 1e-9 * abs(ref_nll)
    assert rel_err(grad.numpy(), ref_grad) <= 1e-9
    gp.factorize(torch.from_numpy(theta))
    side = torch.cuda.Stream()
    base = torch.rand(M, D, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    with torch.cuda.stream(side):
        Xc = base
        for _ in range(40):                       # a long producer chain on the side stream
            Xc = (Xc * 1.000001).remainder(1.0)
        mean, var = gp.predict(Xc)                # ordered after the chain by gpb_set_stream, not by a host sync
    torch.cuda.synchronize()
    assert mean.is_cuda and var.is_cuda and tuple(mean.shape) == (M, 1) and tuple(var.shape) == (M,)
    rmean, rvar = oracle.predict(X, y, Xc.cpu().numpy(), oracle.matern52, theta[:D], theta[D], theta[D + 1])
    assert np.allclose(mean.cpu().numpy(), rmean, rtol=1e-8, atol=1e-12)
    assert np.allclose(var.cpu().numpy(), rvar.ravel(), rtol=1e-8, atol=1e-14)
    K, dK = gp.kernel_matrix(Xc[:300], Xc[:200], torch.from_numpy(theta[:D]), theta[D], theta[D + 1], eval_gradient=True)
    rK, rdK = oracle.matern52(Xc[:300].cpu().numpy(), Xc[:200].cpu().numpy(), theta[:D], theta[D], theta[D + 1],
                              eval_gradient=True)
    assert K.is_cuda and np.allclose(K.cpu().numpy(), rK, rtol=1e-13, atol=1e-15)
    assert np.allclose(dK.cpu().numpy(), rdK, rtol=1e-12, atol=1e-14)
    # a non-SPD covariance raises torch.linalg.LinAlgError (np.linalg.cholesky raises LinAlgError at gp_functions.py:143)
    Xs = torch.from_numpy(np.repeat(np.random.default_rng(0).random((10, 2)), 3, axis=0))
    gs = TorchGP(0, "RBF")
    gs.set_train(Xs, torch.sin(Xs[:, 0]))
    with pytest.raises(torch.linalg.LinAlgError):
        gs.nll(torch.tensor([0.5, 1.0, 0.0], dtype=torch.float64))
    assert gp.launch_count() > 0
    gs.close()
    gp.close()
