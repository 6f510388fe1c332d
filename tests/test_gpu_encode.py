"""GPU encoder parity (-m gpu): the CUDA encoder (model pass + reverse rANS pass + mux/CRC pass), called through the
C ABI, must produce byte-identical .divans streams to the CPU oracle's encoder on the same commands and options."""
import io

import numpy as np
import pytest

from irfuzz import random_ir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def text():
    from divans_b200 import synth
    return synth.text_corpus(1 << 20)


def _first_diff(a, b):
    return next((k for k in range(min(len(a), len(b))) if a[k] != b[k]), min(len(a), len(b)))


def _check(got, ref, what):
    assert len(got) == len(ref) and got == ref, "%s: len %d vs %d, first diff at %d" % (what, len(got), len(ref), _first_diff(got, ref))


OPTION_SETS = [
    dict(),
    dict(window_size=10),
    dict(window_size=16, dynamic_context_mixing=1),
    dict(dynamic_context_mixing=2),
    dict(dynamic_context_mixing=3, prior_depth=2),
    dict(force_stride=0),
    dict(force_stride=3, use_context_map=1),
    dict(use_context_map=0, dynamic_context_mixing=2),
    dict(literal_adaptation=[(2, 1024), (16, 8192), (128, 16384), (1, 128)]),
    dict(dynamic_context_mixing=2, literal_adaptation=[(32, 4096), (4, 1024), (64, 16384), (512, 16384)]),
]


@pytest.mark.parametrize("kw", OPTION_SETS, ids=lambda d: ",".join("%s=%s" % (k, v if not isinstance(v, list) else "set") for k, v in d.items()) or "default")
def test_raw_encode_matches_oracle(engine, oracle, text, kw):
    import divans_b200
    rng = np.random.default_rng(5)
    raws = [text[:n] for n in [0, 1, 2, 15, 16, 17, 1023, 1024, 1025, 4097, 40000]]
    raws.append(rng.integers(0, 256, 3000).astype(np.uint8).tobytes())
    raws.append(bytes(7000))
    got = engine.encode(raws, divans_b200.encode_options(**kw))
    for i, r in enumerate(raws):
        _check(got[i], oracle.encode_raw(r, oracle.options(**kw)), "raw stream %d (len %d) opts %s" % (i, len(r), kw))


def test_raw_encode_chunk_restart_and_big_records(engine, oracle, text):
    # > 65536 literal nibbles (rANS chunk restart, ans.rs:57,138) and coder payloads > 65536 B (fixed-size mux records)
    rng = np.random.default_rng(9)
    raws = [text[:65535], text[:65536], text[:65537], text[:200000], rng.integers(0, 256, 150000).astype(np.uint8).tobytes(),
            text[:32768] + rng.integers(0, 256, 140000).astype(np.uint8).tobytes()]
    got = engine.encode(raws)
    for i, r in enumerate(raws):
        _check(got[i], oracle.encode_raw(r), "stream %d" % i)


@pytest.mark.parametrize("pm", [0, 1, 2, 3])
def test_raw_encode_prediction_modes(engine, oracle, text, pm):
    import divans_b200
    for mv in range(9):
        r = text[5000 * mv: 5000 * mv + 4000]
        out, off, ln = oracle.encode_batch(np.frombuffer(r, np.uint8), [0], [len(r)], oracle.options(), 1, False, pm, mv)
        got = engine.encode([r], divans_b200.encode_options(literal_pred_mode=pm, literal_mixing_value=mv))
        _check(got[0], out[: int(ln[0])].tobytes(), "pm %d mixing %d" % (pm, mv))


def test_command_lists_lz77(engine, oracle, text):
    import divans_b200
    for win, pm, mv, kw in [(16, 2, 4, {}), (10, 0, 4, dict(window_size=10)), (16, 3, 1, dict(dynamic_context_mixing=2)),
                            (16, 1, 6, dict(force_stride=2))]:
        cls = [oracle.Commands.lz77(text[o:o + n], win, pm, mv) for o, n in [(0, 100), (500, 5000), (9000, 70000), (100000, 30000)]]
        kw = dict(kw); kw.setdefault("window_size", win)
        got = engine.encode([c.serialize() for c in cls], divans_b200.encode_options(**kw), cmds=True)
        for i, c in enumerate(cls):
            _check(got[i], c.encode(oracle.options(**kw)), "lz77 list %d win %d" % (i, win))


def test_command_lists_random_ir(engine, oracle, text):
    import divans_b200
    cls, refs = [], []
    for seed in range(24):
        c = oracle.Commands.from_ir(random_ir(oracle, seed, n_cmds=150, window=16, text=text))
        cls.append(c)
    for kw in [dict(window_size=16), dict(window_size=16, dynamic_context_mixing=2, prior_depth=1), dict(window_size=16, force_stride=5)]:
        got = engine.encode([c.serialize() for c in cls], divans_b200.encode_options(**kw), cmds=True)
        for i, c in enumerate(cls):
            _check(got[i], c.encode(oracle.options(**kw)), "random IR seed %d opts %s" % (i, kw))


def test_reference_held_stream_is_reproduced_by_the_gpu_encoder(engine, oracle):
    """Commands of the reference-held stream (wasm/wasm.html:98-107) -> GPU encoder under model revision WASM_2018 ->
    the reference encoder's own 113 bytes; under today's revision -> the oracle's bytes for today's model."""
    import os
    import divans_b200
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    vec = open(os.path.join(d, "ref_wasm_example.divans"), "rb").read()
    rc, plain, cmds = oracle.decode_cmds(vec, model_rev=oracle.MODEL_WASM_2018)
    assert rc == 0
    blob = cmds.serialize()
    kw = dict(window_size=22, use_context_map=0, dynamic_context_mixing=0)
    got = engine.encode([blob, blob], divans_b200.encode_options(model_rev=divans_b200.MODEL_WASM_2018, **kw), cmds=True)
    assert got[0] == vec and got[1] == vec
    now = engine.encode([blob], divans_b200.encode_options(**kw), cmds=True)[0]
    _check(now, cmds.encode(oracle.options(**kw)), "2018 commands under today's model")


def test_bad_command_list_is_rejected_not_crashing(engine):
    blob = np.zeros(64, np.uint8)
    out = np.zeros(1 << 16, np.uint8)
    ln, st = engine.encode_batch_host(blob, [0], [64], out, [0], [1 << 16], None, cmds=True)
    assert st[0] != 0


def test_output_capacity_too_small(engine, oracle, text):
    r = text[:5000]
    ref = oracle.encode_raw(r)
    out = np.zeros(1 << 16, np.uint8)
    blob = np.frombuffer(r, np.uint8)
    ln, st = engine.encode_batch_host(blob, [0], [len(r)], out, [0], [len(ref) - 1])
    assert st[0] == 2 and int(ln[0]) == len(ref)      # NEEDS_MORE_OUTPUT + the size that would have been needed
    ln, st = engine.encode_batch_host(blob, [0], [len(r)], out, [0], [len(ref)])
    assert st[0] == 0 and out[: len(ref)].tobytes() == ref


def test_reference_ffi_compressor_writer(engine, oracle, text):
    import divans_b200
    r = text[:100000]
    sink = io.BytesIO()
    w = divans_b200.DivansCompressorWriter(sink)
    for o in range(0, len(r), 7777):
        w.write(r[o:o + 7777])
    w.close()
    stream = sink.getvalue()
    rc, back = oracle.decode(stream, out_cap=len(r) + 64)
    assert rc == 0 and back == r
    (st, out), = engine.decode([stream], [len(r) + 64])
    assert st == 0 and out == r


def test_full_size_encode_decode_round_trip(engine, oracle):
    # BASELINE config 4 shape: 4096 x 64 KiB through the GPU encoder, back through the GPU decoder
    import divans_b200
    from divans_b200 import synth
    n, size = 4096, 65536
    corpus = synth.text_corpus(n * size // 4)
    raws = [corpus[(i * size) % (len(corpus) - size): (i * size) % (len(corpus) - size) + size] for i in range(n)]
    streams = engine.encode(raws, divans_b200.encode_options(window_size=16))
    for i in [0, 1, n // 2, n - 1]:
        _check(streams[i], oracle.encode_raw(raws[i], oracle.options(window_size=16)), "stream %d" % i)
    res = engine.decode(streams, [size + 64] * n)
    assert all(st == 0 for st, _ in res)
    assert all(out == raws[i] for i, (_, out) in enumerate(res))


def test_ir_text_end_to_end(engine, oracle, text):
    # reference IR text -> product parser (C ABI) -> GPU encoder == oracle encoder; GPU decode == oracle replay of the IR
    import divans_b200
    irs = [random_ir(oracle, 100 + seed, n_cmds=120, window=16, text=text) for seed in range(8)]
    blobs = [divans_b200.ir_to_cmds(ir)[0] for ir in irs]
    got = engine.encode(blobs, divans_b200.encode_options(window_size=16, dynamic_context_mixing=1), cmds=True)
    raws = []
    for i, ir in enumerate(irs):
        c = oracle.Commands.from_ir(ir)
        _check(got[i], c.encode(oracle.options(window_size=16, dynamic_context_mixing=1)), "IR %d" % i)
        rc, raw = c.recode(16)
        assert rc == 0
        raws.append(raw)
    res = engine.decode(got, [len(r) + 64 for r in raws])
    assert all(st == 0 and out == r for (st, out), r in zip(res, raws))


def test_corrupted_command_lists_never_hang(engine, oracle, text):
    # hostile DVCL blobs (random words overwritten in the header / command records): a status per stream, no fault
    rng = np.random.default_rng(5)
    blobs = [np.frombuffer(oracle.Commands.from_ir(random_ir(oracle, 300 + s, n_cmds=100, window=16, text=text)).serialize(), np.uint8) for s in range(10)]
    for _ in range(6):
        bad = []
        for b in blobs:
            w = b.copy().view(np.uint32) if b.size % 4 == 0 else np.concatenate([b, np.zeros(4 - b.size % 4, np.uint8)]).view(np.uint32)
            w = w.copy()
            n_cmds = int(w[2])
            for _m in range(int(rng.integers(1, 5))):
                i = int(rng.integers(2, 8 + 5 * n_cmds))
                w[i] = rng.integers(0, 1 << 32, dtype=np.uint64).astype(np.uint32) if rng.random() < 0.5 else np.uint32(rng.integers(0, 70000))
            bad.append(w.view(np.uint8))
        in_len = np.array([x.size for x in bad], np.uint64)
        in_off = np.zeros(len(bad), np.uint64)
        in_off[1:] = np.cumsum((in_len + np.uint64(15)) & ~np.uint64(15))[:-1]
        blob = np.zeros(int(in_off[-1] + in_len[-1]) + 16, np.uint8)
        for x, o in zip(bad, in_off):
            blob[int(o):int(o) + x.size] = x
        cap = np.full(len(bad), 1 << 20, np.uint64)
        out_off = np.arange(len(bad), dtype=np.uint64) * np.uint64(1 << 20)
        out = np.zeros(len(bad) << 20, np.uint8)
        try:
            ln, st = engine.encode_batch_host(blob, in_off, in_len, out, out_off, cap, None, cmds=True)
            assert all(int(x) in (0, 1, 2, 3) for x in st)
        except Exception as e:      # the host marshaller may refuse a batch whose header asks for absurd sizes
            assert "too large" in str(e) or "encode_batch_host" in str(e)
    good = engine.encode([blobs[0].tobytes()], None, cmds=True)
    assert len(good[0]) > 24
