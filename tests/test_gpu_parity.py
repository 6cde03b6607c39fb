"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes), against the float64
stand-in oracle on the same float32 samples.  STAND-IN PARITY — not EM-Spec output
(/root/reference/README.md:73)."""
import os

import numpy as np
import pytest

import reassign_oracle as orc
from parity_util import check_grid, check_grid_dense, check_grid_rows, check_index, check_points, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SR = 48000
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def emspec(lib_built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import emspec as m
    m.load()
    return m


def run_points(m, x, prm, **kw):
    eng = m.Engine(n_fft=prm.n_fft, hop=prm.hop, noise_gate_db=prm.noise_gate_db,
                   flags=prm.flags | m.FLAG_SYNC, **kw)
    pts = eng.process_points(torch.from_numpy(x).cuda())
    out = tuple(p[0].cpu().numpy() for p in pts)
    eng.close()
    return out


def run_grid(m, x, prm, want_grid=True):
    eng = m.Engine(n_fft=prm.n_fft, hop=prm.hop, noise_gate_db=prm.noise_gate_db,
                   db_range=prm.db_range, gain=prm.gain, low_end_boost=prm.low_end_boost,
                   smoothing=prm.smoothing, sample_rate=prm.sample_rate,
                   display_rows=prm.display_rows, freq_scale=prm.freq_scale,
                   agc_strength=prm.agc_strength, brightness=prm.brightness, flags=prm.flags | m.FLAG_SYNC)
    g, i = eng.process_grid(torch.from_numpy(x).cuda(), want_grid=want_grid)
    eng.close()
    return (g[0].cpu().numpy() if g is not None else None), i[0].cpu().numpy()


@pytest.mark.parametrize("n_fft,hop,secs", [
    (256, 64, 0.25), (512, 128, 0.5), (1024, 256, 0.5), (2048, 512, 2.0),   # configs[0] geometry
    (4096, 128, 1.0), (4096, 256, 1.0), (8192, 256, 1.0), (16384, 4096, 2.0), (32768, 8192, 3.0),
    (4096, 1000, 0.5), (2048, 2048, 0.5),                                   # odd hop, hop = n_fft
])
def test_points_vs_oracle(emspec, n_fft, hop, secs):
    x = orc.synth_signal(int(secs * SR), SR, seed=1)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    stats = check_points(run_points(emspec, x, prm), x, prm)
    assert stats["n_valid"] > 0


def test_points_low_gate_many_valid(emspec):
    """Gate at -120 dB keeps almost every bin: the p99 criterion over a dense point set."""
    x = orc.synth_signal(SR // 2, SR, seed=2)
    prm = orc.Params(n_fft=4096, hop=128, noise_gate_db=-100.0)
    stats = check_points(run_points(emspec, x, prm), x, prm)
    assert stats["n_valid"] > 100000


def test_kat_tone_impulse_chirp(emspec):
    """The analytic KATs of SURVEY.md §4 through the CUDA path."""
    t = np.arange(SR) / SR
    prm = orc.Params(n_fft=2048, hop=512, noise_gate_db=-200.0)
    x = np.sin(2 * np.pi * 1000.37 * t).astype(np.float32)
    dt, dk, e = run_points(emspec, x, prm)
    k = int(np.argmax(e[10]))
    assert abs((k + dk[10, k]) * SR / 2048 - 1000.37) < 2e-3          # wrong sign gives 1015.26
    x = np.zeros(SR, np.float32)
    m, pos = 5, 5 * 512 + 1024 + 300
    x[pos] = 1.0
    dt, dk, e = run_points(emspec, x, prm)
    that = (m + dt[m, 100:900]) * 512 + 1024
    assert np.abs(that - pos).max() < 0.05                             # wrong sign gives pos - 600
    prm = orc.Params(n_fft=4096, hop=128, noise_gate_db=-200.0)
    f0, f1 = 500.0, 8000.0
    x = np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t)).astype(np.float32)
    dt, dk, e = run_points(emspec, x, prm)
    m = 150
    k = int(np.argmax(e[m]))
    for kk in range(k - 3, k + 4):
        tt = ((m + dt[m, kk]) * 128 + 2048) / SR
        assert abs((kk + dk[m, kk]) * SR / 4096 - (f0 + (f1 - f0) * tt)) < 0.05


def test_silence_dc_square_short(emspec):
    prm = orc.Params(n_fft=1024, hop=256)
    for x in (np.zeros(8192, np.float32), np.ones(8192, np.float32),
              np.sign(np.sin(2 * np.pi * 997.0 * np.arange(8192) / SR)).astype(np.float32)):
        pts = run_points(emspec, x, prm)
        assert all(np.isfinite(p).all() for p in pts)
        check_points(pts, x, prm)
    dt, dk, e = run_points(emspec, np.zeros(8192, np.float32), prm)
    assert e.max() == 0.0 and np.abs(dt).max() == 0.0
    g, idx = run_grid(emspec, np.zeros(8192, np.float32), prm)
    assert g.max() == 0.0 and idx.max() == 0
    # shorter than one frame -> zero frames, no launch, no error
    eng = emspec.Engine(n_fft=1024, hop=256)
    assert eng.frame_count(1023) == 0 and eng.frame_count(1024) == 1 and eng.frame_count(1279) == 1
    pts = eng.process_points(torch.zeros(1000, device="cuda"))
    assert pts[0].shape == (1, 0, 513)
    eng.close()


def test_plain_mode(emspec):
    """EMS_FLAG_REASSIGN off: plain |X_h|^2 columns, zero displacement."""
    x = orc.synth_signal(SR // 2, SR, seed=3)
    prm = orc.Params(n_fft=2048, hop=256, flags=orc.FLAG_DETERMINISTIC)
    dt, dk, e = run_points(emspec, x, prm)
    assert np.abs(dt).max() == 0 and np.abs(dk).max() == 0
    check_points((dt, dk, e), x, prm)
    g, _ = run_grid(emspec, x, prm)
    check_grid(g, x, prm)


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (4096, 128), (1024, 64)])
def test_grid_and_index_vs_oracle(emspec, n_fft, hop):
    x = orc.synth_signal(SR, SR, seed=4)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    g, idx = run_grid(emspec, x, prm)
    err, grid_o, _ = check_grid(g, x, prm)
    check_index(idx, grid_o, prm, x)
    # energy conservation: the grid holds exactly the kept point energy
    dt, dk, e = run_points(emspec, x, prm)
    assert abs(g.sum(dtype=np.float64) - e.sum(dtype=np.float64)) <= 1e-5 * e.sum(dtype=np.float64)


@pytest.mark.parametrize("n_fft", [256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_nfft_sweep_display_controls(emspec, n_fft):
    """configs[4]: n_fft sweep 256-32768 at hop = n_fft/4 with low-end boost + smoothing + gate."""
    x = orc.synth_signal(max(SR // 2, 6 * n_fft), SR, seed=12)
    prm = orc.Params(n_fft=n_fft, hop=n_fft // 4, low_end_boost=3.9, smoothing=0.5, noise_gate_db=-65.0)
    check_points(run_points(emspec, x, prm), x, prm)
    g, idx = run_grid(emspec, x, prm)
    err, grid_o, _ = check_grid(g, x, prm)
    check_index(idx, grid_o, prm)


@pytest.mark.parametrize("n_fft,hop,rows,scale", [
    (4096, 128, 546, 1.0), (2048, 512, 256, 0.5), (4096, 256, 1080, 0.0), (1024, 128, 64, 1.5),
])
def test_frequency_scale_display_rows(emspec, n_fft, hop, rows, scale):
    """energy scattered straight onto display rows of the warped frequency
    axis ("Frequency Scale", README.md:48) — grid, index, scatter-from-points and streaming."""
    x = orc.synth_signal(SR, SR, seed=13)
    prm = orc.Params(n_fft=n_fft, hop=hop, display_rows=rows, freq_scale=scale, smoothing=0.0)
    g, idx = run_grid(emspec, x, prm)
    assert g.shape[1] == rows
    # a point sitting on a row boundary of the warped axis may land one row over in fp32: the
    # grid is judged by transport distance; the colour index is judged against the ORACLE's grid,
    # cells fed by a point within tolerance of a row / column rounding boundary excused (<= 1e-4)
    _, grid_o = check_grid_rows(g, x, prm)
    check_index(idx, grid_o, prm, x)
    eng = emspec.Engine(n_fft=n_fft, hop=hop, display_rows=rows, freq_scale=scale,
                        flags=prm.flags | emspec.FLAG_SYNC)
    xd = torch.from_numpy(x).cuda()
    pts = eng.process_points(xd)
    assert pts[0].shape[2] == n_fft // 2 + 1               # points stay per bin
    g2, i2 = eng.scatter_points(*pts)
    assert torch.equal(g2[0].cpu(), torch.from_numpy(g)) and torch.equal(i2[0].cpu(), torch.from_numpy(idx))
    # streaming emits the same rows
    col = torch.empty((1, rows), dtype=torch.uint8).pin_memory()
    S = (len(x) // hop) * hop
    n_ok = 0
    for i in range(S // hop):
        ready, ci = eng.stream_push(torch.from_numpy(x[i * hop:(i + 1) * hop]).contiguous(), col)
        if ready and ci < idx.shape[0]:
            assert (col.numpy()[0] == idx[ci]).all()
            n_ok += 1
    assert n_ok > 10
    eng.close()


@pytest.mark.parametrize("smoothing,strength,rows,brightness", [(0.0, 1.0, 0, 0.44), (0.6, 0.5, 0, 1.0), (0.0, 0.7, 300, 0.7)])
def test_auto_gain_control(emspec, smoothing, strength, rows, brightness):
    """AGC (README.md:14) — column peak, max-with-release level recurrence,
    cells drawn at E / level^strength; offline (sparse and EMA post-pass), host-chunked and
    streaming paths agree with the oracle / with each other."""
    t = np.arange(2 * SR) / SR
    env = 0.02 + 0.9 * (np.sin(2 * np.pi * 0.7 * t) > 0)            # loud / quiet alternation
    x = (env * (0.5 * np.sin(2 * np.pi * 1234.5 * t) + 0.2 * np.sin(2 * np.pi * 5000.25 * t))).astype(np.float32)
    x += orc.synth_signal(2 * SR, SR, seed=17) * 0.05
    prm = orc.Params(n_fft=2048, hop=128, smoothing=smoothing, agc_strength=strength,
                     display_rows=rows, gain=1.0, db_range=50.0, brightness=brightness)
    g, idx = run_grid(emspec, x, prm)
    grid_o, _ = orc.process(x, prm)
    check_index(idx, grid_o, prm, x)                                   # against the ORACLE's grid
    off = orc.postpass(grid_o, orc.Params(**{**prm.__dict__, "agc_strength": 0.0}))
    assert (idx.astype(int) != off.astype(int)).mean() > 1e-3          # the AGC really changes the picture
    if strength == 1.0:    # "Brightness": at full strength the loudest cell of a loud column sits at 255 * brightness
        assert abs(int(idx.max()) - round(255 * brightness)) <= 1
    eng = emspec.Engine(n_fft=2048, hop=128, smoothing=smoothing, agc_strength=strength, display_rows=rows,
                        gain=1.0, db_range=50.0, brightness=brightness, flags=prm.flags | emspec.FLAG_SYNC)
    _, i_host = eng.process_host(torch.from_numpy(x).pin_memory())
    assert (np.abs(i_host[0].numpy().astype(int) - idx.astype(int)) <= 1).all()
    R = idx.shape[1]
    col = torch.empty((1, R), dtype=torch.uint8).pin_memory()
    worst = 0
    n = 0
    for i in range(len(x) // 128):
        ready, ci = eng.stream_push(torch.from_numpy(x[i * 128:(i + 1) * 128]).contiguous(), col)
        if ready and ci < idx.shape[0]:
            worst = max(worst, np.abs(col.numpy()[0].astype(int) - idx[ci].astype(int)).max())
            n += 1
    assert n > 100 and worst <= 1, (n, worst)
    eng.close()


def test_display_controls(emspec):
    """configs[4]: low-end boost + smoothing + noise gate enabled."""
    x = orc.synth_signal(SR, SR, seed=5)
    prm = orc.Params(n_fft=2048, hop=512, low_end_boost=3.9, smoothing=0.5,
                     noise_gate_db=-50.0, db_range=70.0, gain=2.0)
    g, idx = run_grid(emspec, x, prm)
    err, grid_o, _ = check_grid(g, x, prm)
    check_index(idx, grid_o, prm)
    prm.smoothing = 0.9     # long memory: exercises the chunk carries (F > 256 columns)
    prm.hop = 64
    g, idx = run_grid(emspec, x, prm)
    err, grid_o, _ = check_grid(g, x, prm)
    check_index(idx, grid_o, prm)


def test_deterministic_and_fast_mode(emspec):
    x = torch.from_numpy(orc.synth_signal(2 * SR, SR, seed=6)).cuda()
    eng = emspec.Engine(n_fft=4096, hop=128, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    runs = [eng.process_grid(x) for _ in range(3)]
    for g, i in runs[1:]:
        assert torch.equal(g, runs[0][0]) and torch.equal(i, runs[0][1])     # bit-exact
    pts = eng.process_points(x)
    g2, i2 = eng.scatter_points(*pts)                                         # a4 from stored points
    assert torch.equal(g2, runs[0][0]) and torch.equal(i2, runs[0][1])
    eng.close()
    fast = emspec.Engine(n_fft=4096, hop=128, flags=emspec.FLAG_REASSIGN | emspec.FLAG_SYNC)
    gf, _ = fast.process_grid(x)
    fast.close()
    assert rel_l2(gf.cpu().numpy(), runs[0][0].cpu().numpy()) < 1e-5


def test_fixed_point_accumulator_saturates_instead_of_wrapping(emspec):
    """PCM far outside [-1, 1]: a point contributes at most 2^12 in deterministic mode, the
    64-bit sum never wraps (the fp32 fast mode shows what an unbounded sum would be)."""
    t = np.arange(SR // 2) / SR
    x = torch.from_numpy((300.0 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)).cuda()
    det = emspec.Engine(n_fft=2048, hop=64, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    fast = emspec.Engine(n_fft=2048, hop=64, flags=emspec.FLAG_REASSIGN | emspec.FLAG_SYNC)
    gd, _ = det.process_grid(x)
    gf, _ = fast.process_grid(x)
    det.close(); fast.close()
    assert gf.max() > 4096.0 * 4                                   # really beyond the range
    assert torch.isfinite(gd).all() and (gd <= gf * (1 + 1e-5) + 1e-6).all()
    assert gd.max() >= 4096.0                                      # saturated, not wrapped to small values


def test_stereo_planar(emspec):
    xl = orc.synth_signal(SR // 2, SR, seed=7)
    xr = orc.synth_signal(SR // 2, SR, seed=8)
    prm = orc.Params(n_fft=2048, hop=256)
    eng = emspec.Engine(n_fft=2048, hop=256, channels=2, flags=prm.flags | emspec.FLAG_SYNC)
    pcm = torch.from_numpy(np.stack([xl, xr])).cuda()
    pts = eng.process_points(pcm)
    g, idx = eng.process_grid(pcm)
    eng.close()
    for c, x in enumerate((xl, xr)):
        check_points(tuple(p[c].cpu().numpy() for p in pts), x, prm)
        check_grid(g[c].cpu().numpy(), x, prm)


def test_batch_of_clips_as_channels(emspec):
    """configs[3] shape in miniature: equal-length clips passed as planar channels of one call
    give, clip by clip, what single-clip calls give (clips are independent)."""
    n_clips, S = 7, SR // 2
    clips = np.stack([orc.synth_signal(S, SR, clip_index=c) for c in range(n_clips)])
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    batch = emspec.Engine(n_fft=4096, hop=256, channels=n_clips, flags=fl)
    gb, ib = batch.process_grid(torch.from_numpy(clips).cuda())
    pb = batch.process_points(torch.from_numpy(clips).cuda())
    batch.close()
    one = emspec.Engine(n_fft=4096, hop=256, flags=fl)
    for c in range(n_clips):
        g1, i1 = one.process_grid(torch.from_numpy(clips[c]).cuda())
        p1 = one.process_points(torch.from_numpy(clips[c]).cuda())
        assert torch.equal(g1[0], gb[c]) and torch.equal(i1[0], ib[c])
        assert all(torch.equal(a[0], b[c]) for a, b in zip(p1, pb))
    one.close()


def test_process_host_matches_device(emspec):
    """The HOST-buffer call (chunked, overlapped copies) equals the device-buffer call."""
    x = orc.synth_signal(3 * SR, SR, seed=9)
    eng = emspec.Engine(n_fft=1024, hop=32, smoothing=0.3, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    g_dev, i_dev = eng.process_grid(torch.from_numpy(x).cuda())
    g_host, i_host = eng.process_host(torch.from_numpy(x).pin_memory(), want_grid=True)
    assert torch.equal(g_host, g_dev.cpu()) and torch.equal(i_host, i_dev.cpu())
    eng.close()


def test_int16_interleaved_ingest(emspec):
    """int16 interleaved PCM in (capture format) equals the fp32 planar path on
    the same quantised samples, and the oracle on them."""
    xl = orc.synth_signal(SR, SR, seed=14)
    xr = orc.synth_signal(SR, SR, seed=15)
    q = np.clip(np.rint(np.stack([xl, xr], 1) * 32768.0), -32768, 32767).astype(np.int16)   # [S][2]
    xf = (q.astype(np.float32) / 32768.0).T.copy()                                           # [2][S]
    eng = emspec.Engine(n_fft=2048, hop=128, channels=2, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    g16, i16 = eng.process_host_i16(torch.from_numpy(q).pin_memory(), want_grid=True)
    gf, i_f = eng.process_host(torch.from_numpy(xf).pin_memory(), want_grid=True)
    assert torch.equal(g16, gf) and torch.equal(i16, i_f)
    eng.close()
    prm = orc.Params(n_fft=2048, hop=128)
    for c in range(2):
        check_grid(g16[c].numpy(), xf[c], prm)


def test_streaming_on_callers_stream_and_live_controls(emspec):
    """Streaming works when the handle runs on torch's (legacy default) stream, and
    ems_update_display takes effect between pushes, including a change of accumulator type."""
    hop, n_fft = 256, 2048
    x = orc.synth_signal(SR // 2, SR, seed=21)
    eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC)
    eng.use_torch_stream()
    col = torch.empty((1, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    ref = emspec.Engine(n_fft=n_fft, hop=hop, gain=7.0, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    _, want = ref.process_grid(torch.from_numpy(x).cuda(), want_grid=False)
    want = want[0].cpu().numpy()
    ref.close()
    eng.update_display(gain=7.0)
    n = 0
    for i in range(len(x) // hop):
        if i == 40:
            eng.update_display(db_range=58.0)                       # re-captures the graph mid-stream
        ready, ci = eng.stream_push(torch.from_numpy(x[i * hop:(i + 1) * hop]).contiguous(), col)
        if ready and ci < want.shape[0]:
            assert (col.numpy()[0] == want[ci]).all()
            n += 1
    assert n > 50
    eng.update_display(flags=emspec.FLAG_REASSIGN)                   # fp32 accumulator: streaming state restarts
    ready, ci = eng.stream_push(torch.from_numpy(x[:hop]).contiguous(), col)
    assert not ready and ci < 0
    eng.close()


def test_stream_checkpoint_resume(emspec):
    """SURVEY.md §5 checkpoint/resume: a saved stream continues bit-identically in a new handle."""
    hop, n_fft = 256, 2048
    x = orc.synth_signal(SR // 2, SR, seed=22)
    kw = dict(n_fft=n_fft, hop=hop, smoothing=0.3, agc_strength=0.5)
    a = emspec.Engine(**kw)
    col = torch.empty((1, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    hops = [torch.from_numpy(x[i * hop:(i + 1) * hop]).contiguous() for i in range(len(x) // hop)]
    for hp in hops[:37]:
        a.stream_push(hp, col)
    blob = a.stream_save()
    rest_a = []
    for hp in hops[37:]:
        ready, ci = a.stream_push(hp, col)
        rest_a.append((ready, ci, col.numpy().copy()))
    b = emspec.Engine(**kw)
    b.stream_load(blob)
    for hp, (ready_a, ci_a, col_a) in zip(hops[37:], rest_a):
        ready, ci = b.stream_push(hp, col)
        assert (ready, ci) == (ready_a, ci_a) and (not ready or (col.numpy() == col_a).all())
    c = emspec.Engine(n_fft=n_fft, hop=hop // 2)
    with pytest.raises(emspec.EmspecError):
        c.stream_load(blob)                                # different geometry
    for e in (a, b, c):
        e.close()


def test_colour_map_lut_is_bit_exact(emspec):
    """256-entry RGBA table lookup (README.md:15,45), integer work -> bit-exact,
    any alignment and length (head / 16-byte body / tail paths)."""
    rng = np.random.default_rng(5)
    lut = rng.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    eng = emspec.Engine(n_fft=1024, hop=256, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    base = torch.from_numpy(rng.integers(0, 256, 100003, dtype=np.uint8)).cuda()
    for off, n in ((0, 100003), (1, 100000), (7, 15), (13, 4099), (16, 64), (3, 0)):
        idx = base[off:off + n]
        out = eng.colorize(idx.contiguous() if n == 0 else idx, lut)
        want = lut[idx.cpu().numpy()].astype(np.uint32)
        assert (out.cpu().numpy().view(np.uint32) == want).all()
    # on a real image
    x = torch.from_numpy(orc.synth_signal(SR // 2, SR, seed=16)).cuda()
    _, img = eng.process_grid(x, want_grid=False)
    rgba = eng.colorize(img, lut)
    assert (rgba.cpu().numpy().view(np.uint32) == lut[img.cpu().numpy()]).all()
    eng.close()


@pytest.mark.parametrize("hop", [125, 129, 333])
def test_tuned_kernel_odd_hops(emspec, hop):
    """n_fft = 4096 tuned kernel with hops that are odd / not multiples of 4 (tile copies are
    4-byte cp.async, frames start at any sample)."""
    x = orc.synth_signal(SR // 2, SR, seed=18)
    prm = orc.Params(n_fft=4096, hop=hop)
    check_points(run_points(emspec, x, prm), x, prm)


@pytest.mark.parametrize("n_fft,hop", [(256, 17), (256, 64), (256, 256), (512, 33), (512, 100), (512, 512),
                                       (1024, 100), (1024, 257), (1024, 1024), (2048, 96), (2048, 515), (2048, 1400),
                                       (8192, 333), (8192, 8192), (16384, 1000), (32768, 777), (32768, 4096)])
def test_tuned_family_hops(emspec, n_fft, hop):
    """The radix-R x 16 x 16 kernels away from 4096: hops that are multiples of 4 (16-byte tile
    copies), odd (4-byte copies) and so large that a tile holds fewer frames than workers; frame
    counts that leave sub-warp workers (n_fft = 256, 512) with empty slots at the end of a tile."""
    x = orc.synth_signal(max(SR // 2, 3 * n_fft), SR, seed=19)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    check_points(run_points(emspec, x, prm), x, prm)


@pytest.mark.parametrize("n_fft", [256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_tuned_family_matches_generic_and_unaligned_channels(emspec, n_fft, monkeypatch):
    """Stereo with an odd sample count: channel 1 starts off a 16-byte boundary, so one channel
    takes the 16-byte and the other the 4-byte tile copies.  Both must give the oracle's points,
    and the deterministic grid must be bit-identical to the generic kernel's decisions up to
    fp32 rounding of the energies (same cells: checked through the u8 image within +-1)."""
    S = max(SR // 2, 4 * n_fft) + 1
    x = np.stack([orc.synth_signal(S, SR, seed=20), orc.synth_signal(S, SR, seed=21)])
    prm = orc.Params(n_fft=n_fft, hop=n_fft // 8)
    eng = emspec.Engine(n_fft=n_fft, hop=prm.hop, channels=2, flags=prm.flags | emspec.FLAG_SYNC)
    pts = eng.process_points(torch.from_numpy(x).cuda())
    for ch in range(2):
        check_points(tuple(p[ch].cpu().numpy() for p in pts), x[ch], prm)
    _, idx = eng.process_grid(torch.from_numpy(x).cuda(), want_grid=False)
    eng.close()
    monkeypatch.setenv("EMS_FORCE_GENERIC", "1")
    eng = emspec.Engine(n_fft=n_fft, hop=prm.hop, channels=2, flags=prm.flags | emspec.FLAG_SYNC)
    _, idx_g = eng.process_grid(torch.from_numpy(x).cuda(), want_grid=False)
    eng.close()
    a, b = idx.cpu().numpy().astype(np.int16), idx_g.cpu().numpy().astype(np.int16)
    assert np.mean(np.abs(a - b) > 1) < 1e-3


def test_host_chunking_with_tuned_kernel(emspec, monkeypatch):
    """ems_process_host in many small chunks (frame ranges f_begin..f_end of the tuned kernel,
    per-chunk post-pass, EMA and AGC carries across chunks) equals the one-shot device call."""
    monkeypatch.setenv("EMS_HOST_CHUNK_FRAMES", "1100")
    x = orc.synth_signal(12 * SR, SR, seed=19)
    eng = emspec.Engine(n_fft=4096, hop=128, channels=2, smoothing=0.4, agc_strength=0.8,
                        flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    pcm = torch.from_numpy(np.stack([x, x[::-1].copy()]))
    g_dev, i_dev = eng.process_grid(pcm.cuda())
    g_host, i_host = eng.process_host(pcm.pin_memory(), want_grid=True)
    assert torch.equal(g_host, g_dev.cpu())
    d = (i_host.int() - i_dev.cpu().int()).abs()
    assert d.max() <= 1 and (d > 0).float().mean() < 1e-4
    eng.update_display(smoothing=0.0, agc_strength=0.0)
    g2, i2 = eng.process_host(pcm.pin_memory(), want_grid=True)
    g3, i3 = eng.process_grid(pcm.cuda())
    assert torch.equal(g2, g3.cpu()) and torch.equal(i2, i3.cpu())
    eng.close()


def test_golden_fixture(emspec):
    """Committed fixture (tests/golden/make_golden.py): CUDA path vs stored oracle output."""
    z = np.load(os.path.join(GOLD, "reassign_n512_h128.npz"))
    prm = orc.Params(n_fft=int(z["n_fft"]), hop=int(z["hop"]), noise_gate_db=float(z["gate_db"]))
    dt, dk, e = run_points(emspec, z["x"], prm)
    keep = (z["energy"] > 0) & (e > 0)
    assert keep.sum() >= 0.999 * (z["energy"] > 0).sum()
    strong = keep & (z["raw"] >= z["raw"].max() * 1e-4)
    assert np.abs(dt - z["dt_cols"])[strong].max() <= 1e-3
    assert np.abs(dk - z["dk_bins"])[strong].max() <= 1e-3
    g, idx = run_grid(emspec, z["x"], prm)
    assert rel_l2(g, z["grid"]) <= 1e-4
    d = np.abs(idx.astype(int) - z["index"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-3


def test_long_stream_properties(emspec):
    """Full-size geometry (n_fft=4096, hop=128) on a 2-minute stream: properties that do not
    need the oracle — energy conservation, determinism, late frames equal to a shifted run."""
    S = 120 * SR
    g = torch.Generator(device="cuda").manual_seed(0)
    t = torch.arange(S, device="cuda", dtype=torch.float64) / SR
    x = (0.4 * torch.sin(2 * np.pi * (200.0 * t + 30.0 * t * t)) + 0.2 * torch.sin(2 * np.pi * 997.0 * t)).float()
    x += 1e-3 * torch.randn(S, device="cuda", generator=g)
    eng = emspec.Engine(n_fft=4096, hop=128, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    dt, dk, e = eng.process_points(x)
    grid, idx = eng.process_grid(x)
    assert torch.isfinite(dt).all() and torch.isfinite(dk).all() and torch.isfinite(e).all()
    assert dt.abs().max() <= 16.0 + 1e-3                      # |dt| <= N/2 samples = R columns
    tot_p, tot_g = e.double().sum().item(), grid.double().sum().item()
    assert abs(tot_p - tot_g) <= 1e-6 * tot_p
    # shift invariance: frames of x[off*hop:] equal frames off.. of x (same samples, same kernel)
    off = 40000
    dt2, dk2, e2 = eng.process_points(x[off * 128:].contiguous())
    F2 = e2.shape[1]
    inner = slice(20, F2 - 20)     # away from both stream ends (column-in-stream rule)
    assert torch.equal(e2[0, inner], e[0, off:off + F2][inner])
    assert torch.equal(dt2[0, inner], dt[0, off:off + F2][inner])
    # oracle on a late slice (absolute frame index ~ 40000: fp32 displacement form keeps precision)
    sl = x[off * 128: off * 128 + 4096 + 128 * 63].cpu().numpy()
    prm = orc.Params(n_fft=4096, hop=128)
    o = orc.reassign_points(sl, prm)
    a = tuple(v[0, off:off + 64].cpu().numpy() for v in (dt, dk, e))
    both = (o[2] > 0) & (a[2] > 0)
    col_ok = np.ones_like(both)
    col_ok[:17] = False
    col_ok[-17:] = False            # the slice's own stream ends differ from the long stream's
    both &= col_ok
    assert both.sum() > 100
    assert np.percentile(np.abs(a[0] - o[0])[both], 99) <= 1e-3
    assert np.percentile(np.abs(a[1] - o[1])[both], 99) <= 1e-3
    eng.close()


@pytest.mark.parametrize("n_fft,hop,channels,smoothing", [
    (2048, 256, 2, 0.0), (4096, 128, 1, 0.0), (1024, 300, 1, 0.0), (8192, 256, 2, 0.4),
    (256, 64, 2, 0.0), (512, 100, 1, 0.3), (16384, 2048, 1, 0.0), (32768, 4096, 1, 0.0),
])
def test_streaming_matches_offline(emspec, n_fft, hop, channels, smoothing):
    """ems_stream_push (one frame per push, CUDA graph) emits, R pushes late, exactly the columns
    the offline call computes for the same samples (configs[1] geometry is the last case)."""
    S = int(max(0.6 * SR, 2.5 * n_fft)) // hop * hop
    x = np.stack([orc.synth_signal(S, SR, seed=20 + c) for c in range(channels)])
    eng = emspec.Engine(n_fft=n_fft, hop=hop, channels=channels, smoothing=smoothing,
                        flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    _, idx_off = eng.process_grid(torch.from_numpy(x).cuda(), want_grid=False)
    idx_off = idx_off.cpu().numpy()
    B = n_fft // 2 + 1
    col = torch.empty((channels, B), dtype=torch.uint8).pin_memory()
    inter = torch.from_numpy(np.ascontiguousarray(x.T))          # [S][channels] interleaved
    got = {}
    for i in range(S // hop):
        ready, ci = eng.stream_push(inter[i * hop:(i + 1) * hop].contiguous(), col)
        if ready:
            got[ci] = col.numpy().copy()
    F = idx_off.shape[1]
    R = -(-(n_fft // 2) // hop)
    assert sorted(got) == list(range(0, F - R)), (len(got), F, R)
    worst = 0
    for ci, c in got.items():
        d = np.abs(c.astype(int) - idx_off[:, ci].astype(int))
        worst = max(worst, d.max())
    assert worst <= (1 if smoothing > 0 else 0), worst
    # reset: the same pushes give the same columns again
    eng.stream_reset()
    again = {}
    for i in range(S // hop):
        ready, ci = eng.stream_push(inter[i * hop:(i + 1) * hop].contiguous(), col)
        if ready:
            again[ci] = col.numpy().copy()
    assert all((again[k] == got[k]).all() for k in got)
    eng.close()


def test_fused_deposit_never_leaves_the_grid_on_white_noise(emspec):
    """on white noise a few points per 100k frames have k + dk within fp32 rounding of
    N/2 + 0.5; the old test on wh let them through to row N/2 + 1, i.e. row 0 of the next column
    (or past the end of the accumulator).  The fused deposit must equal the bounds-checked scatter of
    the stored points bit for bit, every stored point must land on the grid, and the accumulator must
    be clean afterwards (a second call gives the same image)."""
    S = 4096 + 128 * 160000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = 0.25 * torch.randn(S, device="cuda", generator=g)
    eng = emspec.Engine(n_fft=4096, hop=128, noise_gate_db=-200.0,
                        flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    _, idx = eng.process_grid(x, want_grid=False)
    dt, dk, e = eng.process_points(x)
    k = torch.arange(2049, device="cuda", dtype=torch.float32)[None, None, :]
    row = k + torch.round(dk)                       # torch.round is half-to-even, like rintf
    kept = e > 0
    assert kept.float().mean() > 0.9
    assert bool(((row >= 0) & (row <= 2048))[kept].all())
    near = kept & ((k + dk) > 2048.25)
    assert int(near.sum()) > 0                      # the edge really is exercised
    del row, near, kept, k
    _, idx2 = eng.scatter_points(dt, dk, e, want_grid=False)
    assert torch.equal(idx, idx2)
    del dt, dk, e, idx2
    _, idx3 = eng.process_grid(x, want_grid=False)
    assert torch.equal(idx, idx3)
    eng.close()


@pytest.mark.parametrize("n_fft,hop", [(4096, 128), (8192, 256)])
def test_dense_music_like_signal(emspec, n_fft, hop):
    """default tolerances on a dense, music-like signal (pink noise + harmonic stacks with
    vibrato + noise bursts: ~70 % of all bins above the gate) through points, grid and colour index at
    the geometries of configs[2] and configs[1].  The grid is compared on the cells not fed by a point
    within tolerance of a deposit rounding boundary (nearest-cell deposit is discontinuous there)."""
    x = orc.synth_music(int(1.5 * SR), SR, seed=31)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    stats = check_points(run_points(emspec, x, prm), x, prm)
    assert stats["n_valid"] > 0.4 * stats_total(x, prm)
    g, idx = run_grid(emspec, x, prm)
    err, grid_o, amb = check_grid_dense(g, x, prm)
    check_index(idx, grid_o, prm, x)
    # the same through the warped display axis and through the host path
    prm_w = orc.Params(n_fft=n_fft, hop=hop, display_rows=546)
    g_w, idx_w = run_grid(emspec, x, prm_w)
    _, grid_ow, _ = check_grid_dense(g_w, x, prm_w)
    # rows of the warped axis are a fraction of a bin wide at the low end: 5 % of the cells of this
    # signal are fed by a boundary point, and 1.2e-4 of them actually change colour (measured on B200)
    check_index(idx_w, grid_ow, prm_w, x, max_excused=5e-4)


def stats_total(x, prm):
    return orc.frame_count(len(x), prm.n_fft, prm.hop) * prm.n_bins


def test_int24_packed_ingest(emspec):
    """SURVEY.md §8f-4 / VERDICT r1 #8: packed little-endian int24 PCM, interleaved, equals the fp32
    planar path on the same quantised samples bit for bit (2^-23 steps are exact in fp32); sample
    counts that leave a partial quad at the end of a chunk; three channels (odd byte strides)."""
    S = SR + 3
    xs = [orc.synth_signal(S, SR, seed=40 + c) for c in range(3)]
    q = np.clip(np.rint(np.stack(xs, 1).astype(np.float64) * 8388608.0), -8388608, 8388607).astype(np.int32)   # [S][3]
    xf = (q.astype(np.float32) / 8388608.0).T.copy()                                                          # [3][S]
    b = q.astype("<i4").view(np.uint8).reshape(S, 3, 4)[:, :, :3].copy()                                      # [S][3][3]
    eng = emspec.Engine(n_fft=2048, hop=128, channels=3, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    g24, i24 = eng.process_host_i24(torch.from_numpy(b).pin_memory(), want_grid=True)
    gf, i_f = eng.process_host(torch.from_numpy(xf).pin_memory(), want_grid=True)
    assert torch.equal(g24, gf) and torch.equal(i24, i_f)
    eng.close()
    check_grid(g24[1].numpy(), xf[1], orc.Params(n_fft=2048, hop=128))


def test_stream_push_int16(emspec):
    """the capture-side push takes int16 hops; columns equal the fp32 pushes of the same
    quantised samples, and a stream may switch formats between pushes."""
    hop, n_fft = 256, 2048
    x = orc.synth_signal(SR // 2, SR, seed=41)
    q = np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / 32768.0
    a = emspec.Engine(n_fft=n_fft, hop=hop)
    b = emspec.Engine(n_fft=n_fft, hop=hop)
    ca = torch.empty((1, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    cb = torch.empty((1, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    n = 0
    for i in range(len(x) // hop):
        ra, ia = a.stream_push(torch.from_numpy(xf[i * hop:(i + 1) * hop]).contiguous(), ca)
        src = torch.from_numpy(q[i * hop:(i + 1) * hop]).contiguous() if (i // 10) % 2 == 0 else \
            torch.from_numpy(xf[i * hop:(i + 1) * hop]).contiguous()          # alternate formats every 10 pushes
        rb, ib = b.stream_push(src, cb)
        assert (ra, ia) == (rb, ib)
        if ra:
            assert (ca.numpy() == cb.numpy()).all()
            n += 1
    assert n > 50
    a.close(); b.close()


def test_stream_colormap_columns(emspec):
    """SURVEY §8f-3 on the streaming path: with a colour map set a push also yields the final column as
    RGBA pixels = lut[colour index]; the map can be replaced and switched off between pushes, survives a
    reset, and the pixels are refused when there is no map or no column."""
    hop, n_fft, ch = 256, 2048, 2
    x = np.stack([orc.synth_signal(SR // 2, SR, seed=43), orc.synth_signal(SR // 2, SR, seed=44)], 1).astype(np.float32)
    rng = np.random.default_rng(5)
    lut1 = rng.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    lut2 = rng.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    eng = emspec.Engine(n_fft=n_fft, hop=hop, channels=ch, agc_strength=0.5, smoothing=0.3)
    ref = emspec.Engine(n_fft=n_fft, hop=hop, channels=ch, agc_strength=0.5, smoothing=0.3)
    col = torch.empty((ch, eng.n_rows), dtype=torch.uint8).pin_memory()
    cref = torch.empty((ch, eng.n_rows), dtype=torch.uint8).pin_memory()
    px = torch.empty((ch, eng.n_rows), dtype=torch.int32)
    with pytest.raises(emspec.EmspecError):
        eng.stream_column_rgba(px)                       # no map
    eng.stream_set_colormap(lut1)
    n = {1: 0, 2: 0, 0: 0}
    for i in range(len(x) // hop):
        hopbuf = torch.from_numpy(x[i * hop:(i + 1) * hop].reshape(-1)).contiguous()
        which = (1, 2, 0)[(i // 12) % 3]
        if i % 12 == 0:
            eng.stream_set_colormap({1: lut1, 2: lut2, 0: None}[which])
        r, ci = eng.stream_push(hopbuf, col)
        r2, ci2 = ref.stream_push(hopbuf, cref)
        assert (r, ci) == (r2, ci2)
        if not r or which == 0:
            with pytest.raises(emspec.EmspecError):
                eng.stream_column_rgba(px)               # no column yet / map off
            if r:
                assert (col.numpy() == cref.numpy()).all()
                n[0] += 1
            continue
        assert (col.numpy() == cref.numpy()).all()       # the index column is unchanged by the map
        eng.stream_column_rgba(px)
        lut = lut1 if which == 1 else lut2
        assert (px.numpy().view(np.uint32) == lut[col.numpy()]).all()
        n[which] += 1
    assert min(n.values()) > 10, n
    eng.stream_set_colormap(lut1)
    eng.stream_reset()
    with pytest.raises(emspec.EmspecError):
        eng.stream_column_rgba(px)                       # reset: no column delivered
    for i in range(2 * n_fft // hop):
        r, _ = eng.stream_push(torch.from_numpy(x[i * hop:(i + 1) * hop].reshape(-1)).contiguous(), col)
    assert r
    eng.stream_column_rgba(px)
    assert (px.numpy().view(np.uint32) == lut1[col.numpy()]).all()
    eng.close(); ref.close()


def test_process_host_scratch_is_bounded(emspec):
    """ems_process_host keeps O(chunk) device memory — the scratch of a 10-minute stream
    equals that of a 1-minute one (beyond the AGC-free minimum nothing scales with the stream), stays
    far below the whole-stream accumulator, and the image equals the device-resident call."""
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    sizes = []
    for secs in (120, 600):
        S = secs * SR
        g = torch.Generator(device="cuda").manual_seed(secs)
        x = (0.3 * torch.sin(2 * np.pi * 440.0 * torch.arange(S, device="cuda", dtype=torch.float64) / SR)).float()
        x += 0.01 * torch.randn(S, device="cuda", generator=g)
        eng = emspec.Engine(n_fft=4096, hop=128, flags=fl)
        _, i_host = eng.process_host(x.cpu().pin_memory())
        sizes.append(eng.scratch_bytes())
        F = i_host.shape[1]
        assert sizes[-1] < 2 * 2 ** 30
        if secs == 600:
            assert sizes[-1] < 0.5 * F * 2049 * 8          # the whole-stream u64 accumulator alone
        if secs == 120:
            _, i_dev = eng.process_grid(x, want_grid=False)
            assert torch.equal(i_host, i_dev.cpu())
        eng.close()
    assert sizes[0] == sizes[1], sizes


def test_image_summary_is_exact(emspec):
    """Batch bookkeeping: per-clip (byte sum, position-weighted sum) of the u8 image —
    integer work, bit-exact against NumPy for channel bases at any alignment, short and long images."""
    rng = np.random.default_rng(9)
    for C_, F_, R_ in ((3, 7, 257), (2, 1000, 2049), (5, 1, 129), (1, 40000, 2049)):
        eng = emspec.Engine(n_fft=2 * (R_ - 1), hop=64, channels=C_, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
        img = rng.integers(0, 256, (C_, F_, R_), dtype=np.uint8)
        got = eng.image_summary(torch.from_numpy(img).cuda()).cpu().numpy().astype(np.uint64)
        flat = img.reshape(C_, -1).astype(np.uint64)
        w = (np.arange(flat.shape[1], dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
        assert (got[:, 0] == flat.sum(1)).all()
        assert (got[:, 1] == (flat * w[None, :]).sum(1)).all()
        eng.close()


def test_host_path_formats_chunked_and_tiny(emspec, monkeypatch):
    """ems_process_host* with many small chunks (ring wrap, staging hand-over, ragged int24 quads at chunk
    starts), three channels, fp32 grid output; and the smallest inputs (one frame, one column)."""
    monkeypatch.setenv("EMS_HOST_CHUNK_FRAMES", "1100")
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    S = 5 * SR + 7
    xs = [orc.synth_signal(S, SR, seed=50 + c) for c in range(3)]
    q = np.clip(np.rint(np.stack(xs, 1).astype(np.float64) * 8388608.0), -8388608, 8388607).astype(np.int32)
    xf = (q.astype(np.float32) / 8388608.0).T.copy()
    b = q.astype("<i4").view(np.uint8).reshape(S, 3, 4)[:, :, :3].copy()
    q16 = np.clip(np.rint(np.stack(xs, 1) * 32768.0), -32768, 32767).astype(np.int16)
    eng = emspec.Engine(n_fft=1024, hop=100, channels=3, flags=fl)       # hop not a multiple of 4
    g24, i24 = eng.process_host_i24(torch.from_numpy(b).pin_memory(), want_grid=True)
    gf, i_f = eng.process_host(torch.from_numpy(xf).pin_memory(), want_grid=True)
    gd, i_d = eng.process_grid(torch.from_numpy(xf).cuda())
    assert i24.shape[1] > 2 * 1100                                       # really several chunks
    assert torch.equal(g24, gf) and torch.equal(i24, i_f)
    assert torch.equal(gf, gd.cpu()) and torch.equal(i_f, i_d.cpu())
    g16, i16 = eng.process_host_i16(torch.from_numpy(q16).pin_memory(), want_grid=True)
    x16 = (q16.astype(np.float32) / 32768.0).T.copy()
    g16d, i16d = eng.process_grid(torch.from_numpy(x16).cuda())
    assert torch.equal(g16, g16d.cpu()) and torch.equal(i16, i16d.cpu())
    eng.close()
    for S1 in (1024, 1024 + 99, 1024 + 100):                             # F = 1, 1, 2
        eng = emspec.Engine(n_fft=1024, hop=100, flags=fl)
        x1 = orc.synth_signal(S1, SR, seed=60)
        g_h, i_h = eng.process_host(torch.from_numpy(x1).pin_memory(), want_grid=True)
        g_d, i_d = eng.process_grid(torch.from_numpy(x1).cuda())
        assert g_h.shape[1] == orc.frame_count(S1, 1024, 100) and torch.equal(g_h, g_d.cpu()) and torch.equal(i_h, i_d.cpu())
        eng.close()


def test_dense_and_sparse_post_pass_agree(emspec):
    """The post-pass picks the sparse (flagged blocks) or the dense (whole columns) kernel on the device
    from the dirty flags; both must give the same picture as the oracle, and a handle must be able to
    alternate between them (flags and accumulator clean after either)."""
    x_s = orc.synth_signal(SR, SR, seed=70)
    x_d = orc.synth_music(SR, SR, seed=71)
    for rows in (0, 300):
        for det in (True, False):
            fl = emspec.FLAG_REASSIGN | emspec.FLAG_SYNC | (emspec.FLAG_DETERMINISTIC if det else 0)
            eng = emspec.Engine(n_fft=2048, hop=128, display_rows=rows, flags=fl)       # default gate: the music-like signal
            prm = orc.Params(n_fft=2048, hop=128, display_rows=rows, flags=orc.FLAG_REASSIGN | orc.FLAG_DETERMINISTIC)   # dirties every block
            seen = []
            for x in (x_d, x_s, x_d):
                g, idx = eng.process_grid(torch.from_numpy(x).cuda())
                seen.append((g.cpu(), idx.cpu()))
            if det:
                assert torch.equal(seen[0][1], seen[2][1]) and torch.equal(seen[0][0], seen[2][0])
            else:        # fp32 reductions: the order of the sums, hence a last bit here and there, differs from run to run
                d = (seen[0][1].int() - seen[2][1].int()).abs()
                assert d.max() <= 1 and (d > 0).float().mean() < 1e-3
            for x, (g, idx) in zip((x_d, x_s), seen[:2]):
                if rows == 0:
                    err, grid_o, _ = check_grid_dense(g[0].numpy(), x, prm)
                else:
                    _, grid_o = check_grid_rows(g[0].numpy(), x, prm)
                check_index(idx[0].numpy(), grid_o, prm, x, max_excused=5e-4)
            eng.close()


def test_many_channels(emspec):
    """A batch maps clips to channels: more than 65535 / NB channels used to overflow the post-pass grid."""
    C_, S = 2500, 2048
    rng = np.random.default_rng(11)
    x = (0.2 * rng.standard_normal((C_, S))).astype(np.float32)
    x[:, :] += 0.5 * np.sin(2 * np.pi * (np.arange(C_)[:, None] % 100 + 20) * np.arange(S)[None, :] / 256.0).astype(np.float32)
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    eng = emspec.Engine(n_fft=256, hop=64, channels=C_, flags=fl)
    g, idx = eng.process_grid(torch.from_numpy(x).cuda())
    summ = eng.image_summary(idx).cpu().numpy()
    eng.close()
    one = emspec.Engine(n_fft=256, hop=64, flags=fl)
    for c in (0, 1, 1984, 1985, 1986, 2499):
        g1, i1 = one.process_grid(torch.from_numpy(x[c]).cuda())
        assert torch.equal(g1[0], g[c]) and torch.equal(i1[0], idx[c])
        assert int(summ[c, 0]) == int(i1.sum(dtype=torch.int64))
    one.close()


@pytest.mark.parametrize("n_fft,hop", [(4096, 128), (4096, 333), (1024, 64), (256, 48), (8192, 512)])
def test_schedule_perturbation_is_bit_exact(emspec, n_fft, hop, monkeypatch):
    """Race proxy (compute-sanitizer is closed on this pool): the tile hand-over between workers (mbarrier +
    release counters, no CTA-wide barrier) must give the same bits whatever the schedule.  The persistent
    grid is capped at 1, 3, 17 and 148 CTAs — different tile-to-CTA maps, different numbers of tiles per
    buffer, workers with no frame in short tiles — and points and image must not change by a bit."""
    x = torch.from_numpy(orc.synth_signal(6 * SR + 11, SR, seed=80)).cuda()
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    ref = None
    for ctas in ("148", "1", "3", "17"):
        monkeypatch.setenv("EMS_MAX_CTAS", ctas)
        eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=fl)
        for rep in range(2):
            pts = eng.process_points(x)
            g, idx = eng.process_grid(x)
            cur = tuple(t.clone() for t in (*pts, g, idx))
            if ref is None:
                ref = cur
            else:
                assert all(torch.equal(a, b) for a, b in zip(cur, ref)), (ctas, rep)
        eng.close()
    monkeypatch.setenv("EMS_MAX_CTAS", "5")
    monkeypatch.setenv("EMS_KERNEL_VARIANT", "64")
    if n_fft == 4096:          # the single-exchange variant: its own arithmetic, the same determinism
        eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=fl)
        a = tuple(t.clone() for t in eng.process_points(x))
        eng.close()
        monkeypatch.setenv("EMS_MAX_CTAS", "148")
        eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=fl)
        b = eng.process_points(x)
        assert all(torch.equal(u, v) for u, v in zip(a, b))
        eng.close()


def test_process_grid_bounded_scratch(emspec):
    """EMS_FLAG_BOUNDED_SCRATCH: ems_process_grid in frame chunks on an accumulator ring gives the same bits
    as the one-launch path (deterministic mode), with smoothing + AGC carried across chunks, for one and for
    several channels, and holds a fraction of the memory."""
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    S = 600 * SR
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.arange(S, device="cuda", dtype=torch.float64) / SR
    x = (0.3 * torch.sin(2 * np.pi * (300.0 * t + 10.0 * t * t)) + 0.1 * torch.sin(2 * np.pi * 2500.0 * t)).float()
    x += 2e-3 * torch.randn(S, device="cuda", generator=g)
    for kw in (dict(), dict(smoothing=0.5, agc_strength=0.6)):
        a = emspec.Engine(n_fft=4096, hop=128, flags=fl, **kw)
        b = emspec.Engine(n_fft=4096, hop=128, flags=fl | emspec.FLAG_BOUNDED_SCRATCH, **kw)
        ga, ia = a.process_grid(x)
        gb, ib = b.process_grid(x)
        assert ia.shape[1] > 24 * 148 * 36                      # really more than one chunk
        assert torch.equal(ga, gb)
        d = (ia.int() - ib.int()).abs()
        assert d.max() <= (1 if kw else 0) and (d > 0).float().mean() < 1e-4
        assert b.scratch_bytes() < 0.7 * a.scratch_bytes() and b.scratch_bytes() < 2.5 * 2 ** 30     # constant in the stream length
        a.close(); b.close()
    del ga, gb, ia, ib
    xs = torch.stack([x[: 20 * SR], x[20 * SR: 40 * SR], x[40 * SR: 60 * SR]]).contiguous()
    a = emspec.Engine(n_fft=2048, hop=64, channels=3, flags=fl)
    b = emspec.Engine(n_fft=2048, hop=64, channels=3, flags=fl | emspec.FLAG_BOUNDED_SCRATCH)
    ga, ia = a.process_grid(xs)
    gb, ib = b.process_grid(xs)
    assert torch.equal(ga, gb) and torch.equal(ia, ib)
    a.close(); b.close()


def test_handle_lifecycle_releases_device_memory(emspec):
    """Create / use / destroy many handles along every path (points, grid, bounded scratch, host formats,
    streaming with checkpoint): the device memory in use returns to where it started."""
    torch.cuda.synchronize()
    x = torch.from_numpy(orc.synth_signal(2 * SR, SR, seed=90)).cuda()
    xh = x.cpu().pin_memory()
    q = (x.cpu() * 32768.0).round().clamp_(-32768, 32767).to(torch.int16)[:, None].contiguous().pin_memory()
    free0 = torch.cuda.mem_get_info()[0]
    for it in range(12):
        n_fft, hop = ((4096, 128), (1024, 100), (8192, 512), (256, 64))[it % 4]
        fl = emspec.FLAG_REASSIGN | emspec.FLAG_SYNC | (emspec.FLAG_DETERMINISTIC if it % 2 else 0) | \
            (emspec.FLAG_BOUNDED_SCRATCH if it % 3 == 0 else 0)
        eng = emspec.Engine(n_fft=n_fft, hop=hop, flags=fl, smoothing=0.3 * (it % 2), agc_strength=0.5 * ((it // 2) % 2),
                            display_rows=(0, 200)[it % 2])
        pts = eng.process_points(x)
        g, i = eng.process_grid(x)
        eng.process_host(xh, want_grid=(it % 2 == 0))
        eng.process_host_i16(q)
        col = torch.empty((1, eng.n_rows), dtype=torch.uint8).pin_memory()
        for k in range(2 * n_fft // hop + 3):
            eng.stream_push(xh[k * hop:(k + 1) * hop].contiguous(), col)
        blob = eng.stream_save()
        eng.stream_load(blob)
        eng.stream_push(xh[:hop].contiguous(), col)
        del pts, g, i
        eng.close()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 * 2 ** 20, (free0, free1)


def test_error_behaviour_with_a_live_handle(emspec):
    """The C-ABI never throws or aborts: wrong calls on a live handle return the documented status,
    leave a message in ems_last_error, and the handle keeps working afterwards."""
    import ctypes as C
    lib = emspec.load()
    fl = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    eng = emspec.Engine(n_fft=1024, hop=256, flags=fl)
    h = eng.h
    x = torch.from_numpy(orc.synth_signal(SR // 4, SR, seed=95)).cuda()
    ref_idx = eng.process_grid(x, want_grid=False)[1].clone()
    F, B = ref_idx.shape[1], eng.n_bins
    nf = C.c_size_t()
    buf = torch.empty((1, F, B), dtype=torch.float32, device="cuda")
    u8 = torch.empty((1, F, B), dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    INV, STATE = emspec.ERR_INVALID_ARG, emspec.ERR_STATE

    def err():
        return lib.ems_last_error(h).decode()

    # null buffers
    assert lib.ems_process_points(h, None, x.numel(), p(buf), p(buf), p(buf), C.byref(nf)) == INV and err()
    assert lib.ems_process_points(h, p(x), x.numel(), None, p(buf), p(buf), C.byref(nf)) == INV
    assert lib.ems_process_grid(h, p(x), x.numel(), None, None, C.byref(nf)) == INV and "NULL" in err().upper()
    assert lib.ems_colorize(h, p(u8), 16, None, p(buf)) == INV
    assert lib.ems_image_summary(h, None, F, p(buf)) == INV
    assert lib.ems_stream_column_rgba(h, None) == INV
    # a stream shorter than one frame is zero frames, not an error
    assert lib.ems_process_points(h, p(x), 1000, p(buf), p(buf), p(buf), C.byref(nf)) == emspec.OK and nf.value == 0
    assert eng.frame_count(1023) == 0 and eng.frame_count(1024) == 1
    # geometry cannot change on a live handle; display controls can
    for bad in (dict(n_fft=2048), dict(hop=128), dict(channels=2), dict(display_rows=100), dict(freq_scale=3.0)):
        q = emspec.Params.from_buffer_copy(eng.params)
        for k, v in bad.items():
            setattr(q, k, v)
        assert lib.ems_update_display(h, C.byref(q)) == INV and "new handle" in err(), bad
    q = emspec.Params.from_buffer_copy(eng.params)
    q.smoothing = 1.5
    assert lib.ems_update_display(h, C.byref(q)) == INV
    # state errors
    ms = C.c_float()
    fresh = emspec.Engine(n_fft=1024, hop=256, flags=fl)
    assert lib.ems_stage_ms(fresh.h, emspec.STAGE_POINTS, C.byref(ms)) == STATE
    assert lib.ems_stage_ms(fresh.h, 99, C.byref(ms)) == INV
    px = torch.empty((1, B), dtype=torch.int32)
    assert lib.ems_stream_column_rgba(fresh.h, p(px)) == STATE and "colour map" in lib.ems_last_error(fresh.h).decode()
    # a checkpoint from another geometry, a truncated one and garbage are refused
    col = torch.empty((1, B), dtype=torch.uint8).pin_memory()
    hopbuf = torch.zeros(256, dtype=torch.float32)
    eng.stream_push(hopbuf, col)
    blob = eng.stream_save()
    other = emspec.Engine(n_fft=2048, hop=256, flags=fl)
    other.stream_push(hopbuf, torch.empty((1, 1025), dtype=torch.uint8).pin_memory())
    assert lib.ems_stream_load(other.h, blob, len(blob)) == INV
    assert lib.ems_stream_load(h, blob, len(blob) - 8) == INV
    assert lib.ems_stream_load(h, b"\0" * len(blob), len(blob)) == INV
    assert lib.ems_stream_load(h, blob, len(blob)) == emspec.OK
    # the colour map is configuration: it survives the accumulator-type change that rebuilds the stream state
    lut = np.arange(256, dtype=np.uint32) * 0x01010101
    eng.stream_set_colormap(lut)
    eng.update_display(flags=fl & ~emspec.FLAG_DETERMINISTIC)
    xs = x.cpu()
    for i in range(12):
        r, _ = eng.stream_push(xs[i * 256:(i + 1) * 256].contiguous(), col)
    assert r
    eng.stream_column_rgba(px)
    assert (px.numpy().view(np.uint32) == lut[col.numpy()]).all()
    eng.update_display(flags=fl)
    # after all of that the handle still computes the same image
    assert torch.equal(eng.process_grid(x, want_grid=False)[1], ref_idx)
    for e in (eng, fresh, other):
        e.close()


@pytest.mark.parametrize("n_fft,hop,secs,music", [(4096, 128, 30, False), (4096, 128, 10, True), (8192, 256, 20, False),
                                                   (1024, 64, 10, True)])
def test_points_vs_c_oracle_long(emspec, n_fft, hop, secs, music):
    """Parity at sizes the NumPy oracle is too slow for: tens of seconds of audio (up to 11 k frames,
    23 M points) against the multi-threaded C restatement of the oracle (oracle/reassign_oracle.c, which
    tests/test_c_oracle.py pins to the NumPy one), with the same banded tolerances as the short cases."""
    import c_oracle as co
    x = orc.synth_music(secs * SR, SR, seed=secs) if music else orc.synth_signal(secs * SR, SR, seed=secs)
    prm = orc.Params(n_fft=n_fft, hop=hop)
    ref = co.reassign_points(x, prm, return_raw=True)
    stats = check_points(run_points(emspec, x, prm), x, prm, ref_points=ref)
    assert stats["n_valid"] > 1000 * secs
    # and the image of the same stream: C oracle grid / index vs the GPU
    grid_o = co.scatter_grid(*ref[:3], prm)
    g, i = run_grid(emspec, x, prm)
    if music:
        from parity_util import check_grid_dense
        check_grid_dense(g, x, prm)
    else:
        assert rel_l2(g, grid_o) <= 1e-4
        check_index(i, grid_o, prm, x)


@pytest.mark.parametrize("variant,n_fft,hop", [(64, 4096, 128), (64, 4096, 333), (32, 8192, 2048), (32, 8192, 256), (32, 8192, 1000)])
def test_experimental_kernel_variants_keep_parity(emspec, monkeypatch, variant, n_fft, hop):
    """The kernels kept in the tree as measured experiments (EMS_KERNEL_VARIANT=64: single-exchange 64 x 64 at
    n_fft 4096; =32: three workers, radix-32 first pass, X in place at 8192 — DESIGN.md §4.8) meet the same
    tolerances as the default kernels, points and image, on the sparse and on the music-like signal."""
    monkeypatch.setenv("EMS_KERNEL_VARIANT", str(variant))
    for music in (False, True):
        x = orc.synth_music(2 * SR, SR, seed=variant) if music else orc.synth_signal(2 * SR, SR, seed=variant)
        prm = orc.Params(n_fft=n_fft, hop=hop)
        stats = check_points(run_points(emspec, x, prm), x, prm)
        assert stats["n_valid"] > 0
        g, i = run_grid(emspec, x, prm)
        if music:
            check_grid_dense(g, x, prm)
        else:
            _, grid_o, _ = check_grid(g, x, prm)
            check_index(i, grid_o, prm, x)


def test_stream_push_int24(emspec):
    """The capture-side push also takes packed little-endian int24 hops (three bytes per sample, interleaved,
    full scale 2^23): columns equal the fp32 pushes of the same quantised samples, stereo, and the three
    formats may alternate within one stream."""
    hop, n_fft, ch = 256, 2048, 2
    x = np.stack([orc.synth_signal(SR // 2, SR, seed=61), orc.synth_music(SR // 2, SR, seed=62)], 1)
    q24 = np.clip(np.rint(x.astype(np.float64) * 8388608.0), -8388608, 8388607).astype(np.int32)
    xf = (q24.astype(np.float64) / 8388608.0).astype(np.float32)            # exact: 24 bits fit the fp32 mantissa
    packed = np.stack([(q24 & 0xFF), ((q24 >> 8) & 0xFF), ((q24 >> 16) & 0xFF)], -1).astype(np.uint8)   # [S][ch][3]
    q16 = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    a = emspec.Engine(n_fft=n_fft, hop=hop, channels=ch)
    b = emspec.Engine(n_fft=n_fft, hop=hop, channels=ch)
    ca = torch.empty((ch, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    cb = torch.empty((ch, n_fft // 2 + 1), dtype=torch.uint8).pin_memory()
    n = 0
    for i in range(len(x) // hop):
        sl = slice(i * hop, (i + 1) * hop)
        kind = (i // 7) % 3                    # int24, fp32 of the same samples, int16 (its own quantisation: fed to both)
        if kind == 2:
            src_a = src_b = torch.from_numpy(q16[sl].reshape(-1)).contiguous()
        else:
            src_a = torch.from_numpy(xf[sl].reshape(-1)).contiguous()
            src_b = torch.from_numpy(packed[sl].reshape(-1)).contiguous() if kind == 0 else src_a
        ra, ia = a.stream_push(src_a, ca)
        rb, ib = b.stream_push(src_b, cb)
        assert (ra, ia) == (rb, ib)
        if ra:
            assert (ca.numpy() == cb.numpy()).all(), i
            n += 1
    assert n > 50
    a.close(); b.close()


@pytest.mark.parametrize("kw", [dict(n_fft=4096, hop=128), dict(n_fft=2048, hop=512, display_rows=546, freq_scale=1.0),
                                dict(n_fft=8192, hop=256, display_rows=300, freq_scale=0.0, sample_rate=44100.0)])
def test_cursor_readout_matches_the_oracle(emspec, kw):
    """README.md:39 "note and frequency information" under the cursor: ems_cursor_info is the inverse of
    the row mapping the scatter uses.  Checked against the oracle's cursor_info / row_frequencies on whole and
    fractional rows, out-of-range rows (clamped), and through the picture itself: the brightest row of a
    tone's image reads back as the tone's frequency and note."""
    prm = orc.Params(**kw)
    eng = emspec.Engine(flags=prm.flags | emspec.FLAG_SYNC, **kw)
    R = eng.n_rows
    assert R == prm.n_rows
    fr = orc.row_frequencies(prm)
    rng = np.random.default_rng(3)
    rows = np.concatenate([np.arange(R, dtype=np.float64)[:: max(1, R // 97)], rng.uniform(0, R - 1, 50), [-5.0, R + 3.5, R - 1]])
    for r in rows:
        col = float(rng.integers(0, 100000))
        got, want = eng.cursor_info(col, r), orc.cursor_info(col, r, prm)
        assert abs(got["freq_hz"] - want["freq_hz"]) <= 1e-9 * max(1.0, want["freq_hz"])
        assert abs(got["time_s"] - want["time_s"]) <= 1e-12 * max(1.0, want["time_s"])
        if abs(abs(want["cents"]) - 50.0) > 1e-6:                    # away from the tie between two notes
            assert got["midi_note"] == want["midi_note"] and got["name"] == want["name"]
            assert abs(got["cents"] - want["cents"]) < 1e-4
        if float(r).is_integer() and 0 <= r < R:
            assert abs(got["freq_hz"] - fr[int(r)]) <= 1e-9 * max(1.0, fr[int(r)])
    assert eng.cursor_info(0, 0)["midi_note"] == -1 and eng.cursor_info(0, 0)["name"] == ""
    # the other direction (axis ticks): equals the oracle, inverts the read-out, and rounds to the row the scatter uses
    for hz in (0.0, 27.5, 440.0, 1000.0, 9999.9, prm.sample_rate / 2, 1e6, -3.0):
        r = eng.hz_to_row(hz)
        assert abs(r - orc.hz_to_row(hz, prm)) <= 1e-9 * max(1.0, r)
        if 0 < hz < prm.sample_rate / 2:
            assert abs(eng.cursor_info(0, r)["freq_hz"] - hz) <= 1e-6 * hz
            k = hz * prm.n_fft / prm.sample_rate
            assert int(np.rint(r)) == int(orc.output_row(np.floor(k), np.float64(k - np.floor(k)), prm))
    # A4 through the picture
    sr = prm.sample_rate
    t = np.arange(int(sr)) / sr
    x = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    _, img = eng.process_grid(torch.from_numpy(x).cuda(), want_grid=False)
    img = img[0].cpu().numpy()
    f = img.shape[0] // 2
    info = eng.cursor_info(f, int(np.argmax(img[f])))
    assert info["name"] == "A4" and info["midi_note"] == 69
    row_hz = max(fr[min(R - 1, int(np.argmax(img[f])) + 1)] - fr[int(np.argmax(img[f]))], sr / prm.n_fft)
    assert abs(info["freq_hz"] - 440.0) <= row_hz
    assert abs(info["time_s"] - (f * prm.hop + prm.n_fft / 2) / sr) < 1e-12
    eng.close()


def test_builtin_colour_map_through_the_device_lookup(emspec):
    """A built-in table (ems_colormap_builtin) fed to ems_colorize gives lut[index] on a real image."""
    eng = emspec.Engine(n_fft=1024, hop=256, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    x = torch.from_numpy(orc.synth_signal(SR // 2, SR, seed=21)).cuda()
    _, img = eng.process_grid(x, want_grid=False)
    for name in emspec.colormap_names():
        lut = emspec.builtin_colormap(name)
        assert (lut == orc.builtin_colormap(name)).all()
        rgba = eng.colorize(img, lut)
        assert (rgba.cpu().numpy().view(np.uint32) == lut[img.cpu().numpy()]).all()
    eng.close()


def _sequential_cell_sums(dt, dk, en, F, B, R, chunk):
    """fp32 grid of the sorted scatter, restated: per chunk of `chunk` points (in point order) every cell's
    energies are added left to right in fp32, then the chunk's sum is added to the cell."""
    col = (np.arange(F)[:, None] + np.rint(dt)).astype(np.int64)
    row = (np.arange(B)[None, :] + np.rint(dk)).astype(np.int64)
    ok = (en > 0) & (col >= 0) & (col < F) & (row >= 0) & (row < R)
    key = np.where(ok, col * R + row, -1).reshape(-1)
    e = en.reshape(-1).astype(np.float32)
    acc = np.zeros(F * R, np.float32)
    for i0 in range(0, key.size, chunk):
        k, v = key[i0:i0 + chunk], e[i0:i0 + chunk]
        order = np.argsort(k, kind="stable")
        k, v = k[order], v[order]
        start = np.flatnonzero(np.r_[True, k[1:] != k[:-1]])
        end = np.r_[start[1:], k.size]
        for a, b in zip(start, end):
            if k[a] < 0:
                continue
            s = np.float32(v[a])
            for j in range(a + 1, b):
                s = np.float32(s + v[j])
            acc[k[a]] = np.float32(acc[k[a]] + s)
    return acc.reshape(F, R)


@pytest.mark.parametrize("chunk", [0, 5000, 2048, 777])
def test_sorted_scatter_is_the_ordered_fp32_sum(emspec, chunk, monkeypatch):
    """EMS_FLAG_SORTED_SCATTER (north_star's sort-by-bin + segmented reduce): the grid is, bit for bit, the
    fp32 sum of each cell's energies in ascending point order — also across chunk seams — and agrees with
    the fixed-point deterministic mode to fp32 round-off."""
    if chunk:
        monkeypatch.setenv("EMS_SORT_CHUNK", str(chunk))
    n_fft, hop = 1024, 256
    x = orc.synth_signal(SR // 4, SR, seed=31)
    det = emspec.Engine(n_fft=n_fft, hop=hop, noise_gate_db=-200.0, flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC)
    srt = emspec.Engine(n_fft=n_fft, hop=hop, noise_gate_db=-200.0,
                        flags=emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SORTED_SCATTER | emspec.FLAG_SYNC)
    pts = det.process_points(torch.from_numpy(x).cuda())
    gd, idd = det.scatter_points(*pts)
    gs, ids = srt.scatter_points(*pts)
    gs2, ids2 = srt.scatter_points(*pts)
    assert torch.equal(gs, gs2) and torch.equal(ids, ids2)
    dt, dk, en = (p[0].cpu().numpy() for p in pts)
    F, B = en.shape
    assert (en > 0).mean() > 0.9                                               # dense: cells collide
    want = _sequential_cell_sums(dt, dk, en, F, B, B, chunk or dt.size)
    got = gs[0].cpu().numpy()
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert rel_l2(got, gd[0].cpu().numpy()) < 1e-6
    assert (np.abs(ids[0].cpu().numpy().astype(int) - idd[0].cpu().numpy().astype(int)) <= 1).all()
    det.close(); srt.close()


@pytest.mark.parametrize("kw", [dict(n_fft=4096, hop=128), dict(n_fft=2048, hop=512, channels=2),
                                dict(n_fft=2048, hop=256, display_rows=546, freq_scale=1.0, smoothing=0.4, agc_strength=0.5)])
def test_sorted_scatter_matches_the_default_modes(emspec, kw, monkeypatch):
    """Sorted scatter against the fixed-point mode on sparse and multi-channel input, the warped display
    axis with smoothing and AGC behind it, several chunks, and a handle that alternates between calls."""
    monkeypatch.setenv("EMS_SORT_CHUNK", str(300000))
    ch = kw.get("channels", 1)
    x = np.stack([orc.synth_signal(SR, SR, seed=50 + c) for c in range(ch)])
    base = emspec.FLAG_REASSIGN | emspec.FLAG_DETERMINISTIC | emspec.FLAG_SYNC
    det = emspec.Engine(flags=base, **kw)
    srt = emspec.Engine(flags=base | emspec.FLAG_SORTED_SCATTER, **kw)
    xd = torch.from_numpy(x).cuda()
    pts = det.process_points(xd)
    gd, idd = det.scatter_points(*pts)
    for _ in range(2):
        gs, ids = srt.scatter_points(*pts)
        assert rel_l2(gs.cpu().numpy(), gd.cpu().numpy()) < 1e-6
        d = np.abs(ids.cpu().numpy().astype(int) - idd.cpu().numpy().astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
        g2, i2 = srt.process_grid(xd)                                          # the fused path ignores the flag
        g3, i3 = det.process_grid(xd)
        assert torch.equal(g2, g3) and torch.equal(i2, i3)
    det.close(); srt.close()
