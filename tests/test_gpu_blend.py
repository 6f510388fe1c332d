"""GPU parity for the reference's feature="blend" probability model (-m gpu): decoder and encoder kernels
instantiated with BlendCDF16 (divans_b200/csrc/dv_blend.cuh), called through the C ABI (DIVANS_B200_FLAG_CDF_BLEND /
divans_b200_encode_options::cdf_model), against the oracle's blend build (oracle/oracle_blend.py) on the same inputs."""
import numpy as np
import pytest

from irfuzz import random_ir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def text():
    from divans_b200 import synth
    return synth.text_corpus(1 << 20)


def _raws(text):
    rng = np.random.default_rng(11)
    return [text[:n] for n in [0, 1, 2, 15, 16, 17, 1023, 1025, 4097, 40000, 65536, 150000]] + [
        rng.integers(0, 256, 3000).astype(np.uint8).tobytes(), bytes(7000), (b"abcabcabd" * 2000)]


def test_decode_matches_blend_oracle(engine, oracle_blend, text):
    import divans_b200
    raws = _raws(text)
    for kw in [dict(), dict(dynamic_context_mixing=2), dict(dynamic_context_mixing=1, prior_depth=1, force_stride=3)]:
        streams = [oracle_blend.encode_raw(r, oracle_blend.options(**kw)) for r in raws]
        res = engine.decode(streams, [len(r) + 64 for r in raws], divans_b200.FLAG_CDF_BLEND)
        for i, (st, out) in enumerate(res):
            assert st == 0 and out == raws[i], (kw, i, st)


def test_decode_prediction_modes_and_mixing_values(engine, oracle_blend, text):
    import divans_b200
    streams, raws = [], []
    for pm in range(4):
        for mv in range(9):
            r = text[1000 * mv:1000 * mv + 3000 + 17 * pm]
            out, off, ln = oracle_blend.encode_batch(np.frombuffer(r, np.uint8), [0], [len(r)], oracle_blend.options(dynamic_context_mixing=mv % 3),
                                                     1, False, pm, mv)
            streams.append(out[: int(ln[0])].tobytes())
            raws.append(r)
    res = engine.decode(streams, [len(r) + 64 for r in raws], divans_b200.FLAG_CDF_BLEND)
    for i, (st, out) in enumerate(res):
        assert st == 0 and out == raws[i], i


def test_decode_random_ir_and_lz77(engine, oracle_blend, text):
    import divans_b200
    streams, want = [], []
    for seed in range(24):
        win = [10, 14, 16, 22][seed % 4]
        c = oracle_blend.Commands.from_ir(random_ir(oracle_blend, 500 + seed, n_cmds=150, window=win, text=text))
        streams.append(c.encode(oracle_blend.options(window_size=c.window, dynamic_context_mixing=seed % 3, prior_depth=seed % 4,
                                                     use_context_map=0 if seed % 7 == 3 else 1)))
        want.append(c.recode(c.window)[1])
    for n in (5000, 70000):
        c = oracle_blend.Commands.lz77(text[:n], window=16)
        streams.append(c.encode(oracle_blend.options(window_size=16)))
        want.append(text[:n])
    res = engine.decode(streams, [len(w) + 64 for w in want], divans_b200.FLAG_CDF_BLEND)
    for i, (st, out) in enumerate(res):
        rc, ref = oracle_blend.decode(streams[i])
        assert rc == 0 and ref == want[i]
        assert st == 0 and out == ref, i


def test_encode_matches_blend_oracle(engine, oracle_blend, text):
    import divans_b200
    raws = _raws(text)
    for kw in [dict(), dict(dynamic_context_mixing=2, prior_depth=1), dict(window_size=16, force_stride=5)]:
        got = engine.encode(raws, divans_b200.encode_options(cdf_model=divans_b200.CDF_BLEND, **kw))
        for i, r in enumerate(raws):
            assert got[i] == oracle_blend.encode_raw(r, oracle_blend.options(**kw)), (kw, i, len(r))
    cls = [oracle_blend.Commands.from_ir(random_ir(oracle_blend, 900 + seed, n_cmds=120, window=16, text=text)) for seed in range(12)]
    for kw in [dict(window_size=16), dict(window_size=16, dynamic_context_mixing=2)]:
        got = engine.encode([c.serialize() for c in cls], divans_b200.encode_options(cdf_model=divans_b200.CDF_BLEND, **kw), cmds=True)
        for i, c in enumerate(cls):
            assert got[i] == c.encode(oracle_blend.options(**kw)), (kw, i)


def test_models_do_not_mix_and_slots_survive_a_model_change(engine, oracle, oracle_blend, text):
    """Nothing in a stream names its model: a blend stream decoded as frequentist (and vice versa) must behave like the oracle of
    the wrong model does -- and a slot that hosted a blend stream (step counts in the sign bits of its priors) must be clean for
    the tagged priors of the next default-model batch."""
    import divans_b200
    raw = text[:30000]
    sb, sf = oracle_blend.encode_raw(raw), oracle.encode_raw(raw)
    for flags, stream, wrong in ((0, sb, oracle), (divans_b200.FLAG_CDF_BLEND, sf, oracle_blend)):
        st, out = engine.decode([stream], [len(raw) + 64], flags | divans_b200.FLAG_SKIP_CRC)[0]
        rc, ref = wrong.decode(stream, out_cap=len(raw) + 64, skip_crc=True)
        assert st == rc, (flags, st, rc)
        if rc == 0:
            assert out == ref
    for _ in range(2):
        n = 64
        res = engine.decode([sb] * n, [len(raw) + 64] * n, divans_b200.FLAG_CDF_BLEND)
        assert all(st == 0 and out == raw for st, out in res)
        res = engine.decode([sf] * n, [len(raw) + 64] * n)
        assert all(st == 0 and out == raw for st, out in res)


def test_full_batch_round_trip(engine):
    """1024 x 64 KiB through the blend encoder and decoder (size-independent property: decode(encode(x)) == x)."""
    import divans_b200
    from divans_b200 import synth
    blob, off, ln = synth.text_streams(1024, 65536, seed=21)
    raws = [blob[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
    streams = engine.encode(raws, divans_b200.encode_options(cdf_model=divans_b200.CDF_BLEND))
    res = engine.decode(streams, [len(r) + 64 for r in raws], divans_b200.FLAG_CDF_BLEND)
    assert all(st == 0 and out == r for (st, out), r in zip(res, raws))


def test_reference_ffi_as_a_blend_deployment(oracle_blend, text):
    """The reference's FFI has no option for the probability model (it is a cargo feature): a process that replaces a
    `--features blend` build says so through DIVANS_B200_FFI_CDF=blend.  Run in a child process (the switch is read once)."""
    import os
    import subprocess
    import sys
    import tempfile
    raw = text[:50000]
    enc = oracle_blend.encode_raw(raw, oracle_blend.options(dynamic_context_mixing=1))
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "in.divans"), "wb").write(enc)
        open(os.path.join(d, "in.raw"), "wb").write(raw)
        code = (
            "import io, sys\n"
            "sys.path.insert(0, %r)\n"
            "import divans_b200\n"
            "enc = open(%r, 'rb').read(); raw = open(%r, 'rb').read()\n"
            "rd = divans_b200.DivansDecompressorReader(io.BytesIO(enc), 4096)\n"
            "out = bytearray(); chunk = bytearray(4096)\n"
            "while True:\n"
            "    n = rd.readinto(chunk)\n"
            "    if not n: break\n"
            "    out += chunk[:n]\n"
            "rd.close()\n"
            "assert bytes(out) == raw, 'FFI decode of a blend stream'\n"
            "sink = io.BytesIO(); w = divans_b200.DivansCompressorWriter(sink); w.write(raw); w.close()\n"
            "open(%r, 'wb').write(sink.getvalue())\n"
        ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(d, "in.divans"), os.path.join(d, "in.raw"), os.path.join(d, "out.divans"))
        env = dict(os.environ, DIVANS_B200_FFI_CDF="blend")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
        mine = open(os.path.join(d, "out.divans"), "rb").read()
    rc, back = oracle_blend.decode(mine)      # what the FFI compressor wrote is a blend stream the blend oracle reads
    assert rc == 0 and back == raw


def test_truncated_and_corrupt_blend_streams_match_the_oracle(engine, oracle_blend, text):
    import divans_b200
    raw = text[:20000]
    enc = oracle_blend.encode_raw(raw)
    rng = np.random.default_rng(17)
    streams = [enc[:k] for k in (0, 7, 16, 40, len(enc) // 2, len(enc) - 9, len(enc) - 1)]
    for _ in range(24):
        b = bytearray(enc)
        pos = int(rng.integers(16, len(enc) - 8))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(b))
    caps = [len(raw) + 64] * len(streams)
    res = engine.decode(streams, caps, divans_b200.FLAG_CDF_BLEND | divans_b200.FLAG_SKIP_CRC)
    for i, (st, out) in enumerate(res):
        rc, ref = oracle_blend.decode(streams[i], out_cap=caps[i], skip_crc=True)
        assert st == rc, (i, st, rc)
        if rc == 0:
            assert out == ref, i
