"""Host-side layout logic of directional plans (dims=..., src/directional.jl, test/accuracy.jl:83-163): the batched
kernels see (transformed dims..., batch); these helpers move any other arrangement there and back."""
import itertools

import numpy as np
import pytest

import nfft_jl_b200 as nb
from nfft_jl_b200.plan import _dir_from_internal, _dir_to_internal, _normalise_dims


@pytest.mark.parametrize("shape,npre,nlead", [((3, 4, 5), 1, 1), ((3, 4, 5), 0, 1), ((3, 4, 5), 2, 1), ((3, 4, 5), 1, 2),
                                               ((3, 4, 5), 0, 2), ((2, 3, 4, 5), 1, 2), ((6, 7), 1, 1)])
def test_directional_permutation_semantics(shape, npre, nlead):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    pre, lead, post = shape[:npre], shape[npre:npre + nlead], shape[npre + nlead:]
    xi = _dir_to_internal(x, npre, nlead)
    assert xi.shape == lead + (int(np.prod(pre + post)),) and xi.flags.f_contiguous
    b = 0      # batch index runs over (pre..., post...) in column-major order: exactly the slices the reference loops over
    for Ipost in itertools.product(*[range(n) for n in post[::-1]]):
        for Ipre in itertools.product(*[range(n) for n in pre[::-1]]):
            idx = tuple(Ipre[::-1]) + tuple(slice(None) for _ in lead) + tuple(Ipost[::-1])
            assert np.array_equal(xi[..., b], x[idx])
            b += 1
    out = np.zeros(shape, dtype=complex)
    assert np.array_equal(_dir_from_internal(xi, pre, post, out), x)
    xt = _dir_to_internal(torch.from_numpy(x), npre, nlead)
    assert np.array_equal(xt.numpy(), xi) and xt.stride()[0] == 1
    ot = torch.zeros(shape, dtype=torch.complex128)
    assert np.array_equal(_dir_from_internal(xt, pre, post, ot).numpy(), x)


def test_dims_argument():
    assert _normalise_dims(None, 3) == (1, 2, 3)
    assert _normalise_dims(2, 3) == (2,)
    assert _normalise_dims(range(2, 4), 3) == (2, 3)
    for bad in [(1, 3), (0, 1), (3, 4), ()]:
        with pytest.raises(nb.ArgumentError):
            _normalise_dims(bad, 3)
