"""tcgen05 split-bf16 GEMM (aoenv_gemm_tn_tc) against float64 and against the FP32 SIMT kernel."""
import pytest
import torch

from parity_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _run(dev, M, N, K, parts, seed=0, scale_rows=False):
    from rlao_b200 import gemm
    Kp = (K + 15) // 16 * 16
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.zeros(M, Kp, device=dev)
    W = torch.zeros(N, Kp, device=dev)
    X[:, :K] = torch.randn(M, K, device=dev, generator=g)
    W[:, :K] = torch.randn(N, K, device=dev, generator=g)
    if scale_rows:       # wide dynamic range, like DM commands (1e-7 m) against unit influence functions
        X *= 1e-7 * torch.exp(3 * torch.randn(M, 1, device=dev, generator=g))
    op = gemm.Operator(W, parts=parts)
    ldd = (N + 3) // 4 * 4
    D = torch.full((M, ldd), float("nan"), device=dev)
    gemm.gemm_tn(X, op, D, M, N, alpha=0.5, backend="tc")
    torch.cuda.synchronize()
    ref = 0.5 * (X.double() @ W.double().T)
    mag = 0.5 * (X.double().abs() @ W.double().abs().T)      # sum |x||w|: the scale the split error is relative to
    err = ((D[:, :N].double() - ref).abs() / mag).max().item()
    assert torch.isnan(D[:, N:]).all()
    return err


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1, 1, 16), (37, 131, 48), (300, 257, 1360), (1024, 980, 2928), (1000, 1353, 2528)])
def test_tc_gemm_two_parts(dev, M, N, K):
    assert _run(dev, M, N, K, 2) < 2.0 ** -16


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (37, 131, 48), (1024, 980, 2928)])
def test_tc_gemm_three_parts(dev, M, N, K):
    assert _run(dev, M, N, K, 3) < 2.0 ** -18


@pytest.mark.parametrize("M,N,K,parts", [(3072, 980, 2928, 3), (5000, 1300, 200, 2), (2400, 1000, 64, 3)])
def test_tc_gemm_stream_k_order(dev, M, N, K, parts):
    """More output tiles than SMs (192, 220, 152): the kernel cuts tiles x k-blocks into equal ranges per SM, tiles shared
    by two CTAs are summed with float atomics on a zeroed D — accuracy as in tile order, and two runs agree bit for bit
    (three layers of cfg3 extruded together are the first shape)."""
    from rlao_b200 import gemm
    assert _run(dev, M, N, K, parts) < (2.0 ** -16 if parts == 2 else 2.0 ** -18)
    Kp = (K + 15) // 16 * 16
    g = torch.Generator(device=dev).manual_seed(7)
    X, W = torch.randn(M, Kp, device=dev, generator=g), torch.randn(N, Kp, device=dev, generator=g)
    op = gemm.Operator(W, parts=parts)
    ldd = (N + 3) // 4 * 4
    D1, D2 = torch.zeros(M, ldd, device=dev), torch.ones(M, ldd, device=dev)
    gemm.gemm_tn(X, op, D1, M, N, backend="tc")
    gemm.gemm_tn(X, op, D2, M, N, backend="tc")
    assert torch.equal(D1[:, :N], D2[:, :N])


def test_tc_gemm_wide_dynamic_range(dev):
    assert _run(dev, 64, 4096, 368, 2, scale_rows=True) < 2.0 ** -15


def test_tc_matches_simt_on_dm_shape(dev):
    from rlao_b200 import gemm
    g = torch.Generator(device=dev).manual_seed(3)
    M, N, Kp = 96, 57600, 1360
    X = torch.randn(M, Kp, device=dev, generator=g) * 1e-7
    W = torch.rand(N, Kp, device=dev, generator=g)
    op = gemm.Operator(W, parts=2)
    D1 = torch.zeros(M, N, device=dev)
    D2 = torch.zeros(M, N, device=dev)
    gemm.gemm_tn(X, op, D1, M, N, backend="tc")
    gemm.gemm_tn(X, op, D2, M, N, backend="simt")
    assert rel_err(D1.cpu().numpy(), D2.cpu().numpy()) < 2e-5
