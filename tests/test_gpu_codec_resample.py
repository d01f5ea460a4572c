"""GPU parity (through the C ABI): G.711, linear resample, polyphase resample -- all BIT-EXACT."""
import numpy as np
import pytest

from oracle import codec, resample

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab(gpu):
    from open_speech_b200.realtime import audio_buffer

    return audio_buffer


def test_g711_all_codes_and_golden(gpu, ab, golden):
    allb = bytes(range(256))
    assert ab.decode_audio_to_pcm16(allb, "g711_ulaw", 8000) == golden["codec_ulaw_all256_8k"].tobytes()
    assert ab.decode_audio_to_pcm16(allb, "g711_alaw", 8000) == golden["codec_alaw_all256_8k"].tobytes()
    ul, al = golden["codec_ulaw_in"].tobytes(), golden["codec_alaw_in"].tobytes()
    assert ab.decode_audio_to_pcm16(ul, "g711_ulaw", 16000) == golden["codec_ulaw_16k"].tobytes()
    assert ab.decode_audio_to_pcm16(al, "g711_alaw", 16000) == golden["codec_alaw_16k"].tobytes()
    assert ab.decode_audio_to_pcm16(ul[:160], "g711_ulaw", 16000) == golden["codec_ulaw_chunk160_16k"].tobytes()
    assert ab.decode_audio_to_pcm16(golden["codec_pcm24k_in"].tobytes(), "pcm16", 16000) == golden["codec_pcm24k_16k"].tobytes()
    p = golden["codec_pcm16k_in"].tobytes()
    assert ab.encode_pcm16_to_format(p, 16000, "g711_ulaw") == golden["codec_enc_ulaw"].tobytes()
    assert ab.encode_pcm16_to_format(p, 16000, "g711_alaw") == golden["codec_enc_alaw"].tobytes()
    assert ab.encode_pcm16_to_format(p, 16000, "pcm16") == golden["codec_enc_pcm16"].tobytes()


def test_g711_encode_all_65536(gpu, golden):
    all16 = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    for law, key in ((gpu.FMT_ULAW, "codec_lin2ulaw_all"), (gpu.FMT_ALAW, "codec_lin2alaw_all")):
        out = np.empty(65536, np.uint8)
        gpu.call("osb_g711_encode_host", gpu.ptr(all16), gpu.ptr(out), 65536, law)
        assert np.array_equal(out, golden[key])
    # ragged tail (n % 16 != 0)
    out = np.empty(1003, np.uint8)
    part = all16[30000:31003].copy()  # keep a reference: gpu.ptr() is a bare address
    gpu.call("osb_g711_encode_host", gpu.ptr(part), gpu.ptr(out), 1003, gpu.FMT_ULAW)
    assert np.array_equal(out, golden["codec_lin2ulaw_all"][30000:31003])


@pytest.mark.parametrize("n,fr,to", [(160, 8000, 16000), (480, 24000, 16000), (2400, 24000, 16000), (1, 8000, 16000),
                                     (2, 8000, 16000), (7, 16000, 8000), (441, 44100, 16000), (3, 48000, 16000),
                                     (100, 16000, 24000), (99999, 8000, 16000), (1600, 16000, 8000), (5, 48000, 8000)])
def test_linear_resample_bit_exact(gpu, ab, n, fr, to):
    rng = np.random.default_rng(n + fr)
    x = rng.integers(-32768, 32768, n).astype(np.int16)
    assert ab._resample_linear(x.tobytes(), fr, to) == codec.resample_linear(x.tobytes(), fr, to)


@pytest.mark.parametrize("n,m", [(160, 320), (480, 320), (2400, 1600), (320, 160), (441, 160), (161, 321), (7, 13), (1000, 3000), (2, 5)])
def test_linear_resample_integer_fast_path_adversarial(gpu, n, m):
    """The integer fast path of k_resample_linear against np.interp on inputs built to hit its hand-over cases: interpolants that are
    exactly integral (sample differences that are multiples of m-1), grids with common factors (interior coincidences), flat segments,
    full-scale steps, many independent chunks per launch.  Every output must equal numpy's float64 result bit for bit."""
    rng = np.random.default_rng(n * 1000 + m)
    batch = 257
    x = rng.integers(-32768, 32768, (batch, n)).astype(np.int64)
    x[0::5] = (x[0::5] // (m - 1)) * (m - 1)                 # differences divisible by m-1: exactly integral interpolants
    x[1::7] = x[1::7, :1]                                    # flat
    x[2::11] = np.where(rng.integers(0, 2, (len(x[2::11]), n)) > 0, 32767, -32768)  # full-scale steps
    x[3::13] = np.cumsum(rng.integers(-3, 4, (len(x[3::13]), n)), axis=1) * (m - 1) // 2  # half-integral slopes
    x = np.clip(x, -32768, 32767).astype(np.int16)
    out = np.empty((batch, m), np.int16)
    gpu.call("osb_resample_linear_host", gpu.ptr(x), gpu.FMT_PCM16, gpu.ptr(out), gpu.FMT_PCM16, n, m, batch, n, m)
    xo, xn = np.linspace(0, 1, n), np.linspace(0, 1, m)
    for b in range(batch):
        want = np.interp(xn, xo, x[b].astype(np.float32)).astype(np.int16)   # audio_buffer.py:28-33
        assert np.array_equal(out[b], want), (b, np.nonzero(out[b] != want)[0][:5])


@pytest.mark.parametrize("fmt,n,m", [("ulaw", 160, 320), ("alaw", 160, 320), ("pcm16", 160, 320), ("pcm16", 480, 320), ("ulaw", 400, 800), ("pcm16", 882, 320)])
def test_linear_resample_tiled_kernel(gpu, fmt, n, m, monkeypatch):
    """The tiled kernel (many short dense rows: the realtime door replayed over whole recordings) against the per-output kernel on the same
    input, bit for bit, and against np.interp on a sample of rows.  The batch is not a multiple of the tile (32 rows), the rows carry the
    adversarial patterns of the test above, and one row per tile region is all hand-over cases (exactly integral interpolants)."""
    rng = np.random.default_rng(n + m)
    batch = 16384 + 37
    code = {"ulaw": gpu.FMT_ULAW, "alaw": gpu.FMT_ALAW, "pcm16": gpu.FMT_PCM16}[fmt]
    if fmt == "pcm16":
        x = rng.integers(-32768, 32768, (batch, n)).astype(np.int64)
        x[0::5] = (x[0::5] // (m - 1)) * (m - 1)
        x[1::7] = x[1::7, :1]
        x[2::11] = np.where(rng.integers(0, 2, (len(x[2::11]), n)) > 0, 32767, -32768)
        x = np.clip(x, -32768, 32767).astype(np.int16)
    else:
        x = rng.integers(0, 256, (batch, n)).astype(np.uint8)
        x[1::7] = x[1::7, :1]
    tiled = np.empty((batch, m), np.int16)
    gpu.call("osb_resample_linear_host", gpu.ptr(x), code, gpu.ptr(tiled), gpu.FMT_PCM16, n, m, batch, n, m)
    monkeypatch.setenv("OSB_LINEAR_NO_TILES", "1")
    plain = np.empty((batch, m), np.int16)
    gpu.call("osb_resample_linear_host", gpu.ptr(x), code, gpu.ptr(plain), gpu.FMT_PCM16, n, m, batch, n, m)
    monkeypatch.delenv("OSB_LINEAR_NO_TILES")
    assert np.array_equal(tiled, plain), np.argwhere(tiled != plain)[:5]
    xo, xn = np.linspace(0, 1, n), np.linspace(0, 1, m)
    for b in (0, 1, 2, 5, 31, 32, 16383, 16384, batch - 1):
        lin = x[b] if fmt == "pcm16" else np.frombuffer((codec.ulaw2lin if fmt == "ulaw" else codec.alaw2lin)(x[b].tobytes()), np.int16)
        want = np.interp(xn, xo, lin.astype(np.float32)).astype(np.int16)
        assert np.array_equal(tiled[b], want), b


def test_linear_resample_edges(gpu, ab):
    assert ab._resample_linear(b"", 8000, 16000) == b""
    assert ab.decode_audio_to_pcm16(b"", "g711_ulaw", 16000) == b""
    assert ab._resample_linear(np.array([5], np.int16).tobytes(), 48000, 8000) == b""  # out_len == 0
    assert ab.encode_pcm16_to_format(b"", 16000, "g711_ulaw") == b""


def test_realtime_tick_batch_1024_streams(gpu):
    """BASELINE config 3 shape: 1024 streams x 160 mu-law bytes -> 320 pcm16 each, one launch."""
    from open_speech_b200 import synth

    ticks = synth.ulaw_streams(1024, 3)  # [3, 1024, 160]
    for t in range(3):
        inp = np.ascontiguousarray(ticks[t])
        out = np.empty((1024, 320), np.int16)
        gpu.call("osb_resample_linear_host", gpu.ptr(inp), gpu.FMT_ULAW, gpu.ptr(out), gpu.FMT_PCM16, 160, 320, 1024, 160, 320)
        for s in (0, 1, 17, 511, 1023):
            assert out[s].tobytes() == codec.decode_audio_to_pcm16(inp[s].tobytes(), "g711_ulaw", 16000)


@pytest.mark.parametrize("fr", [8000, 24000, 48000, 44100, 22050, 32000])
def test_poly_golden(gpu, golden, fr):
    from open_speech_b200.streaming import resample_pcm16

    assert resample_pcm16(golden[f"poly_{fr}_in"].tobytes(), fr, 16000) == golden[f"poly_{fr}_out"].tobytes()


@pytest.mark.parametrize("n,fr,to", [(160, 8000, 16000), (1600, 8000, 16000), (2, 8000, 16000), (3, 48000, 16000),
                                     (100, 16000, 48000), (441, 44100, 16000), (1601, 16000, 8000), (96000, 8000, 16000),
                                     (4800, 48000, 16000), (100, 16000, 32000), (300, 48000, 16000)])
def test_poly_bit_exact_vs_scipy(gpu, n, fr, to):
    from open_speech_b200.streaming import resample_pcm16

    rng = np.random.default_rng(n + fr + to)
    x = rng.integers(-32768, 32768, n).astype(np.int16)
    got, ref = resample_pcm16(x.tobytes(), fr, to), resample.resample_pcm16(x.tobytes(), fr, to)
    assert len(got) == len(ref)
    assert got == ref


def test_poly_reference_behaviour(gpu, golden):
    """The properties tests/test_streaming_units.py:37-96 pins."""
    from open_speech_b200.streaming import resample_pcm16

    assert len(resample_pcm16(np.arange(100, dtype=np.int16).tobytes(), 16000, 32000)) == 400
    assert len(resample_pcm16(np.arange(300, dtype=np.int16).tobytes(), 48000, 16000)) == 200
    assert len(resample_pcm16(np.arange(441, dtype=np.int16).tobytes(), 44100, 16000)) == 2 * int(441 * 16000 / 44100)
    out = np.frombuffer(resample_pcm16(np.full(100, 5000, np.int16).tobytes(), 16000, 48000), np.int16)
    assert np.allclose(out, 5000, atol=4)
    assert resample_pcm16(golden["poly_up_in"].tobytes(), 16000, 48000) == golden["poly_up_48k_out"].tobytes()
    ext = np.array([32767, -32768, 0, 16000, -16000], dtype=np.int16)
    o = np.frombuffer(resample_pcm16(ext.tobytes(), 8000, 16000), np.int16)
    assert o.tobytes() == resample.resample_pcm16(ext.tobytes(), 8000, 16000)


def test_large_decode_roundtrip_property(gpu):
    """Full-size property check: decode(encode(decode(b))) == decode(b) for every byte (G.711 idempotence)."""
    rng = np.random.default_rng(3)
    n = 1 << 22
    b = rng.integers(0, 256, n).astype(np.uint8)
    for law in (gpu.FMT_ULAW, gpu.FMT_ALAW):
        d1 = np.empty(n, np.int16)
        gpu.call("osb_g711_decode_host", gpu.ptr(b), gpu.ptr(d1), n, law)
        e = np.empty(n, np.uint8)
        gpu.call("osb_g711_encode_host", gpu.ptr(d1), gpu.ptr(e), n, law)
        d2 = np.empty(n, np.int16)
        gpu.call("osb_g711_decode_host", gpu.ptr(e), gpu.ptr(d2), n, law)
        assert np.array_equal(d1, d2)
        tab = codec.ulaw2lin_table() if law == gpu.FMT_ULAW else codec.alaw2lin_table()
        assert np.array_equal(d1, tab[b])


def test_realtime_tick_cuda_graph(gpu):
    """The captured-graph tick (H2D + kernel + D2H in one launch) gives the same bytes as the oracle."""
    import torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import RealtimeTickGraph

    ticks = synth.ulaw_streams(256, 4)
    rg = RealtimeTickGraph(256)
    for t in range(4):
        rg.host_in.copy_(torch.from_numpy(np.ascontiguousarray(ticks[t])))
        out = rg.run().numpy()
        for s in (0, 7, 255):
            assert out[s].tobytes() == codec.decode_audio_to_pcm16(ticks[t, s].tobytes(), "g711_ulaw", 16000)
