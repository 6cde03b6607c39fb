"""Property tests: random geometries and signals.  The CPU half checks
invariants of the stand-in oracle; the GPU half checks the CUDA path against it."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import reassign_oracle as orc

SR = 48000
geom = st.tuples(st.sampled_from([256, 512, 1024, 2048]), st.sampled_from([2, 3, 4, 8, 16, 32]))
kinds = st.sampled_from(["silence", "dc", "square", "noise", "tone", "impulses"])


def make_signal(kind, n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    if kind == "silence":
        x = np.zeros(n)
    elif kind == "dc":
        x = np.full(n, 0.7)
    elif kind == "square":
        x = np.sign(np.sin(2 * np.pi * (50 + 3000 * rng.random()) * t / SR))
    elif kind == "noise":
        x = 0.3 * rng.standard_normal(n)
    elif kind == "tone":
        x = 0.8 * np.sin(2 * np.pi * (20 + 20000 * rng.random()) * t / SR)
    else:
        x = np.zeros(n)
        x[rng.integers(0, n, 5)] = 1.0
    return x.astype(np.float32)


@settings(max_examples=25, derandomize=True, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(geom, kinds, st.integers(0, 2 ** 16), st.integers(0, 700))
def test_oracle_invariants(g, kind, seed, extra):
    n_fft, div = g
    hop = max(1, n_fft // div)
    n = n_fft + 7 * hop + extra                       # first and last partial frames exercised by `extra`
    x = make_signal(kind, n, seed)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    dt, dk, e = orc.reassign_points(x, prm)
    F = orc.frame_count(n, n_fft, hop)
    assert dt.shape == (F, n_fft // 2 + 1)
    assert np.isfinite(dt).all() and np.isfinite(dk).all() and np.isfinite(e).all() and (e >= 0).all()
    assert (np.abs(dt) * hop <= n_fft / 2 + 1e-9).all()
    k = np.arange(n_fft // 2 + 1)[None, :]
    assert ((k + dk >= -0.5) & (k + dk <= n_fft / 2 + 0.5)).all()
    grid = orc.scatter_grid(dt, dk, e, prm)
    assert abs(grid.sum() - e.sum()) <= 1e-9 * max(e.sum(), 1e-30)      # energy conservation
    idx = orc.postpass(grid, prm)
    assert idx.shape == grid.shape and idx.dtype == np.uint8
    if kind == "silence":
        assert grid.max() == 0 and idx.max() == 0


@pytest.mark.gpu
@settings(max_examples=int(__import__("os").environ.get("EMS_HYP_EXAMPLES", "12")), derandomize=True, deadline=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(st.tuples(st.sampled_from([256, 1024, 4096, 8192]), st.sampled_from([2, 4, 16, 32])), kinds,
       st.integers(0, 2 ** 16), st.integers(0, 3000))
def test_gpu_matches_oracle_on_random_cases(lib_built, g, kind, seed, extra):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import emspec
    from parity_util import check_grid, check_points
    n_fft, div = g
    hop = n_fft // div
    n = n_fft + 40 * hop + extra
    x = make_signal(kind, n, seed)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=prm.flags | emspec.FLAG_SYNC)
    xd = torch.from_numpy(x).cuda()
    pts = tuple(p[0].cpu().numpy() for p in eng.process_points(xd))
    grid, _ = eng.process_grid(xd)
    eng.close()
    assert all(np.isfinite(p).all() for p in pts)
    # The spec'd tolerances (1e-3 p99, 1e-4 grid) are stated for the bench signal at 4096/128 and
    # are tested as such in test_gpu_parity.py.  Random cases get the bounds fp32 can honour in
    # general: the p99 population of a bare tone is its -60 dB leakage skirt, whose error grows
    # with n_fft; and the nearest-cell deposit is discontinuous, so on broadband noise the few
    # 1e-4 of points within fp32 error of a rounding boundary land one cell over.
    check_points(pts, x, prm, p99_tol=1e-3 * max(1.0, n_fft / 2048))
    if kind in ("silence", "dc", "tone"):
        check_grid(grid[0].cpu().numpy(), x, prm)
    elif kind == "noise":
        check_grid(grid[0].cpu().numpy(), x, prm, tol=1e-2)
