"""Batch-stationary 2-D kernels (csrc/twod_batch.cuh: Float32, ntransforms >= 8): every transform of a batched plan
against the numpy oracle run on that transform alone, and against the per-transform kernels (kernel_mode 9) of the
same plan.  Covers the three CTA batch sizes (8 / 16 / 32 transforms), batches that do not fill the last CTA, all
footprints the kernels take (m = 2, 3, 4), every window evaluation mode, ragged grids (last tile narrower, periodic
wrap of the halo), explicit tile sizes, and a radial trajectory whose centre tile is split into strided work items."""
import numpy as np
import pytest

from oracle import nfft_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def nb():
    import nfft_jl_b200 as m
    m.lib()
    return m


def check(nb, k, N, m, B, **kw):
    T = np.float32
    p = nb.plan_nfft(np.ascontiguousarray(k.T), N, m=m, σ=2.0, ntransforms=B, **kw)
    okw = {}
    if "precompute" in kw:
        okw["precompute"] = int(kw["precompute"])
    po = O.OraclePlan(k, N, m=m, sigma=2.0, blockSize=p.params.blockSize, **okw)
    assert np.array_equal(p.permutation()[0], po.perm)
    M = k.shape[0]
    f = O.random_complex(tuple(N) + (B,), T, 4)
    fh = O.random_complex((M, B), T, 5)
    fwd = np.asarray(p * f)
    adj = np.asarray(p.adjoint() * fh)
    p.set_kernel_mode(9)                                   # per-transform kernels of the same plan
    fwd9 = np.asarray(p * f)
    adj9 = np.asarray(p.adjoint() * fh)
    assert rel(fwd, fwd9) <= 2e-6 and rel(adj, adj9) <= 2e-6
    p.set_kernel_mode(7)                                   # the batch-stationary kernels without register windows
    assert rel(np.asarray(p.adjoint() * fh), adj9) <= 2e-6
    assert rel(np.asarray(p * f), fwd9) <= 2e-6
    for b in sorted({0, 1, B // 2, B - 1}):
        ef = rel(fwd[:, b], po.forward(np.asfortranarray(f[..., b])))
        ea = rel(adj[..., b], po.adjoint(np.ascontiguousarray(fh[:, b])))
        assert ef <= TOL and ea <= TOL, (b, ef, ea)
    return p


@pytest.mark.parametrize("B", [8, 12, 16, 31, 32, 40])
def test_batch_sizes(nb, B):
    k = O.random_nodes(6000, 2, np.float32, seed=B)
    p = check(nb, k, (48, 40), 4, B)
    assert tuple(p.params.blockSize) == (16, 16)


@pytest.mark.parametrize("m", [2, 3, 4])
@pytest.mark.parametrize("N", [(20, 36), (33, 17), (64, 64)])
def test_footprints_and_ragged_grids(nb, m, N):
    k = O.random_nodes(3000, 2, np.float32, seed=m)
    check(nb, k, N, m, 9)


@pytest.mark.parametrize("pre", ["LINEAR", "FULL", "TENSOR", "POLYNOMIAL"])
def test_window_modes(nb, pre):
    k = O.random_nodes(2500, 2, np.float32, seed=7)
    check(nb, k, (40, 24), 4, 8, precompute=getattr(nb, pre))


@pytest.mark.parametrize("bs", [(8, 8), (16, 8), (32, 32)])
def test_explicit_tiles(nb, bs):
    k = O.random_nodes(5000, 2, np.float32, seed=3)
    check(nb, k, (64, 48), 4, 16, blockSize=bs)


def test_radial_crowded_centre_and_empty_tiles(nb):
    """256 spokes x 256 samples on a 128^2 image: the centre tiles hold thousands of nodes (strided work items), the
    corner tiles none"""
    k = O.radial_nodes(256, 256, np.float32)
    check(nb, k, (128, 128), 4, 32)


def test_single_node_and_tiny_batches_of_nodes(nb):
    for M in (1, 63, 64, 65, 129):
        k = O.random_nodes(M, 2, np.float32, seed=M)
        check(nb, k, (24, 24), 3, 8)
