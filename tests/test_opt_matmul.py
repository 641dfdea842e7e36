"""OptMatmul (deepmd/source/op/opt_matmul.cc:24-62): the oracle's restatement against the reference's own
engine loop (deepmd/source/op/graph.h compiled where it lies, oracle/_ref) on the CPU, and the GPU product (FP64 tensor
cores) against the oracle through the C ABI, tolerance 1e-12 * (|xx| |w|)."""
import numpy as np
import pytest

SHAPES = [(1, 1, 1), (7, 3, 5), (37, 25, 13), (300, 1, 25), (129, 25, 50), (257, 50, 100), (1000, 240, 240), (64, 17, 129)]


@pytest.mark.parametrize("shape", SHAPES[:6])
def test_oracle_is_the_reference_engine_loop(oracle, ref, shape):
    if not ref.available:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    xx, w = rng.uniform(-1, 1, (M, N)), rng.uniform(-1, 1, (N, K))
    a, b = oracle.opt_matmul(xx, w), ref.opt_matmul(xx, w)
    assert np.array_equal(a, b)                      # same k-ascending multiply-then-add per output entry
    assert np.allclose(a, xx @ w, rtol=0, atol=1e-13 * N)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_gpu_opt_matmul_matches_oracle(oracle, shape):
    import g4s_b200

    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    xx, w = rng.uniform(-1, 1, (M, N)), rng.uniform(-1, 1, (N, K))
    got = g4s_b200.opt_matmul(xx, w)
    want = oracle.opt_matmul(xx, w)
    scale = oracle.opt_matmul(np.abs(xx), np.abs(w))
    assert np.all(np.abs(got - want) <= 1e-12 * scale + 1e-300)


@pytest.mark.gpu
def test_gpu_opt_matmul_empty_inner_dimension():
    import g4s_b200

    assert np.array_equal(g4s_b200.opt_matmul(np.zeros((5, 0)), np.zeros((0, 3))), np.zeros((5, 3)))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(37, 25, 13), (500, 100, 240), (129, 3, 1)])
def test_gpu_opt_matmul_gradient_products(oracle, shape):
    """_opt_matmul_grad.py: dxx = grad w^T, dw = xx^T grad — the same kernel with swapped strides."""
    import torch

    from g4s_b200.opt_matmul import opt_matmul_grad_device

    M, N, K = shape
    rng = np.random.default_rng(K)
    xx, w, grad = rng.uniform(-1, 1, (M, N)), rng.uniform(-1, 1, (N, K)), rng.uniform(-1, 1, (M, K))
    t = [torch.from_numpy(a).cuda() for a in (xx, w, grad)]
    dxx = torch.empty(M, N, dtype=torch.float64, device="cuda")
    dw = torch.empty(N, K, dtype=torch.float64, device="cuda")
    opt_matmul_grad_device(M, N, K, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), dxx.data_ptr(), dw.data_ptr())
    torch.cuda.synchronize()
    want_dxx = oracle.opt_matmul(grad, np.ascontiguousarray(w.T))
    want_dw = oracle.opt_matmul(np.ascontiguousarray(xx.T), grad)
    assert np.all(np.abs(dxx.cpu().numpy() - want_dxx) <= 1e-12 * oracle.opt_matmul(np.abs(grad), np.abs(w.T).copy()))
    assert np.all(np.abs(dw.cpu().numpy() - want_dw) <= 1e-12 * oracle.opt_matmul(np.abs(xx.T).copy(), np.abs(grad)))
