"""NRD front-end packing (shaders/includes/rt/nrd_frontend.glsl:11-41, pt_raygen_offline.rgen:106-127)."""
import numpy as np
import pytest

from oracle import oracle_py as O


def _inputs(n=200000, seed=3):
    rng = np.random.default_rng(seed)
    a = np.zeros((n, 6), np.float32)
    a[:, :3] = rng.gamma(0.7, 2.0, (n, 3)); a[: n // 50, :3] *= -1                       # a few negative radiances (clamped to 0 by the packer)
    a[:, 3] = rng.uniform(-1, 200, n); a[:, 4] = rng.uniform(-150, 150, n); a[:, 5] = rng.uniform(0, 1.2, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.concatenate([d, rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    b[:6, :3] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -1, 0], [0.6, 0.8, 0], [-0.6, 0, -0.8]]; b[:6, 3] = [0, 0, 0.5, 1, 0.001, 0.25]
    return a, b


def test_known_answers_and_round_trip():
    a = np.array([[1, 1, 1, 0, 0, 1], [4, 2, 0, 23, 100, 0], [0.2, 0.4, 0.6, 5, -10, 0.5]], np.float32)
    b = np.array([[0, 0, 1, 0.5], [0, 0, -1, 0.5], [1, 0, 0, 0.0]], np.float32)
    pr, pn, back = O.nrd_pack(a, b)
    assert np.allclose(pr[0], [1, 0, 0, 0]) and np.allclose(pr[1], [2, 2, 0, 23.0 / (13.0 * 20.0)])          # Y, Co, Cg; hd / ((3 + 100 * 0.1) * 20)
    assert np.allclose(pr[2, 3], 5.0 / ((3 + 1.0) * (20 * 0.75 + 0.25)))
    assert np.allclose(pn[0], [0.5, 0.5, 0.75, 0]) and np.allclose(pn[1], [0.5, 0.5, 0.25, 0])                 # sign of n.z rides in the roughness channel
    assert np.allclose(pn[2], [1.0, 0.0, 0.5 + 0.5 * 1.5 / 512, 0])                                            # roughness clamped away from 0
    assert np.allclose(back, np.maximum(a[:, :3], 0), atol=1e-6)                                               # YCoCg -> linear inverts the packing


def test_product_code_matches_oracle_on_the_host():
    from tests.emul import emul_py as E
    a, b = _inputs()
    for x, y in zip(O.nrd_pack(a, b), E.nrd_pack(a, b)):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from ohao_engine_b200 import binding as B, scenes
    a, b = _inputs()
    r = B.Renderer(16, 16); r.set_scene(scenes.cornell_box())
    ref, got = O.nrd_pack(a, b), r.nrd_pack(a, b)
    # fp32 mul/add chains, contracted into FMAs on the device: a few ulp of the LARGEST operand (Co / Cg cancel when r ~ b)
    scale = np.abs(a[:, :3]).max(1, keepdims=True) + 1.0
    assert (np.abs(ref[0] - got[0]) <= 5e-7 * scale).all() and (np.abs(ref[2] - got[2]) <= 1e-6 * scale).all()
    assert np.abs(ref[1] - got[1]).max() <= 5e-7
    assert np.array_equal(ref[1][:, 3], got[1][:, 3]) and (np.sign(ref[1][:, 2] - 0.5) == np.sign(got[1][:, 2] - 0.5)).all()
