"""tcgen05 building blocks (pnce_selftest_umma): bulk copy -> smem, UMMA descriptors for the
no-swizzle canonical layout (K-major A/B, MN-major B), commit, tcgen05.ld.  Exact integer-valued
bf16 operands, so the comparison with numpy is exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tile_blob(x):
    """(R, K) -> [K/8][R/8][8 rows][8 k] bf16: 8x(16 B) core matrices, 128 contiguous bytes each."""
    r, k = x.shape
    return x.reshape(r // 8, 8, k // 8, 8).permute(2, 0, 1, 3).contiguous().to(torch.bfloat16)


def run_probe(a_blob, b_blob, a_desc, b_desc, n, k, b_mn):
    from gan_variant_research_b200 import _lib
    lib = _lib.load()
    d = torch.full((128, n), float("nan"), dtype=torch.float32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    a, b = a_blob.cuda(), b_blob.cuda()
    rc = lib.pnce_selftest_umma(a.data_ptr(), a.numel() * 2, b.data_ptr(), b.numel() * 2,
                                a_desc[0], a_desc[1], a_desc[2], b_desc[0], b_desc[1], b_desc[2],
                                n, k, b_mn, d.data_ptr(), err.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "pnce_selftest_umma")
    torch.cuda.synchronize()
    return d.cpu(), int(err.item())


@pytest.fixture(scope="module")
def need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("n,k", [(256, 32), (128, 64), (256, 256)])
def test_kmajor_operands(need_cuda, n, k):
    """D = A B^T as in phase 1 of k_loss_tc (Z = Q K^T)."""
    g = torch.Generator().manual_seed(n + k)
    a = torch.randint(-4, 5, (128, k), generator=g).float()
    b = torch.randint(-4, 5, (n, k), generator=g).float()
    d, err = run_probe(tile_blob(a), tile_blob(b), (16 * 128, 128, 2 * 16 * 128),
                       (n // 8 * 128, 128, 2 * (n // 8) * 128), n, k, 0)
    assert err == 0
    np.testing.assert_array_equal(d.numpy(), (a @ b.t()).numpy())


@pytest.mark.parametrize("n,k", [(32, 256), (32, 128), (64, 256)])
def test_mn_major_b_operand(need_cuda, n, k):
    """D = A Kmat with Kmat (k rows j, n channels c) stored [c/8][j/8][8 j][8 c] -- the very same
    image phase 1 reads K-major -- read MN-major as in phase 2 of k_loss_tc (dQ = dZ K)."""
    g = torch.Generator().manual_seed(7 * n + k)
    a = torch.randint(-4, 5, (128, k), generator=g).float()
    kmat = torch.randint(-4, 5, (k, n), generator=g).float()
    d, err = run_probe(tile_blob(a), tile_blob(kmat), (16 * 128, 128, 2 * 16 * 128),
                       (128, k // 8 * 128, 2 * 128), n, k, 1)
    assert err == 0
    np.testing.assert_array_equal(d.numpy(), (a @ kmat).numpy())


@pytest.mark.parametrize("n,k", [(64, 32), (256, 32), (256, 64), (128, 128)])
def test_mn_major_b_operand_key_major_blob(need_cuda, n, k):
    """The gather's second K blob: [j/8][c/8][8 j][8 c] (a chunk of keys x all channels), read
    MN-major with N = all channels as in phase 2 of k_loss_tc (dQ = dZ K, M128 x N=C x K16)."""
    g = torch.Generator().manual_seed(11 * n + k)
    a = torch.randint(-4, 5, (128, k), generator=g).float()
    kmat = torch.randint(-4, 5, (k, n), generator=g).float()
    blob = kmat.reshape(k // 8, 8, n // 8, 8).permute(0, 2, 1, 3).contiguous().to(torch.bfloat16)
    d, err = run_probe(tile_blob(a), blob, (16 * 128, 128, 2 * 16 * 128),
                       (n // 8 * 128, 128, 2 * (n // 8) * 128), n, k, 1)
    assert err == 0
    np.testing.assert_array_equal(d.numpy(), (a @ kmat).numpy())


@pytest.mark.parametrize("n,k", [(64, 32), (256, 128)])
def test_mn_major_a_and_b_operands(need_cuda, n, k):
    """Weight gradients (k_wgrad_tc): D = A^T-view . B with BOTH operands read MN-major from row blobs
    [col/8][row/8][8 rows][8 cols] whose rows are the contraction index."""
    g = torch.Generator().manual_seed(13 * n + k)
    at = torch.randint(-4, 5, (k, 128), generator=g).float()      # (contraction m, output row i)
    bt = torch.randint(-4, 5, (k, n), generator=g).float()        # (contraction m, output col j)
    d, err = run_probe(tile_blob(at), tile_blob(bt), (128, k // 8 * 128, 2 * 128),
                       (128, k // 8 * 128, 2 * 128), n, k, 3)
    assert err == 0
    np.testing.assert_array_equal(d.numpy(), (at.t() @ bt).numpy())
