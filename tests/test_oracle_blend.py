"""CPU tests of the oracle's feature="blend" build (BlendCDF16, reference src/probability/blend_cdf.rs:109-208).

The reference holds no stream coded with this model; what it does hold are the four tests it runs on BlendCDF16 through
declare_common_tests! (blend_cdf.rs:213, probability/common_tests.rs:4-110) and test_blend_lut (blend_cdf.rs:215-224).  They are
restated here against the oracle's implementation, so the blend model is pinned at property level (divans_oracle.h says so)."""
import ctypes

import numpy as np

CDF_MAX = 32767
MED = (0x30, 0x4000)   # Speed::MED -- the blend model computes a rate from it and does not use it (blend_cdf.rs:186-208)


def _cdf(ob):
    c = ob.Cdf16()
    ob.lib().dvo_cdf_default(ctypes.byref(c))
    return c


def _blend(ob, c, sym):
    ob.lib().dvo_cdf_blend(ctypes.byref(c), sym, ob.Speed(*MED))


def _start_freq(ob, c, sym):
    st, fr = ctypes.c_int16(), ctypes.c_int16()
    ob.lib().dvo_cdf_sym_start_freq(ctypes.byref(c), sym, ctypes.byref(st), ctypes.byref(fr))
    return st.value, fr.value


def test_default_is_the_zero_cdf_and_reads_as_uniform(oracle_blend):
    c = _cdf(oracle_blend)
    assert list(c.c) == [0] * 16 and c.mix_rate == 1536 and c.count == 0           # blend_cdf.rs:128-136
    vals = [oracle_blend.lib().dvo_cdf_value(ctypes.byref(c), s) for s in range(16)]
    assert vals == [(CDF_MAX * (s + 1)) >> 4 for s in range(15)] + [CDF_MAX]         # :160-171 with cdf[15] == 0


def test_sym_to_start_and_freq(oracle_blend):
    # common_tests.rs:4-22: consecutive symbols tile the range, each start = previous end + 1 (the "+1" of interface.rs:103)
    c = _cdf(oracle_blend)
    for i in range(100):
        _blend(oracle_blend, c, i & 0xf)
        last = (0, 0)
        for sym in range(16):
            start, freq = _start_freq(oracle_blend, c, sym)
            assert start == 1 + (0 if sym == 0 else last[0] + last[1])
            last = (start, freq)


def test_cdf_offset_to_sym_start_and_freq(oracle_blend):
    # common_tests.rs:24-43: the symbol found for an offset is monotone in the offset, the offset lies in its range, 15 is reached
    L = oracle_blend.lib()
    c = _cdf(oracle_blend)
    st, fr = ctypes.c_int16(), ctypes.c_int16()
    for i in range(100):
        _blend(oracle_blend, c, i & 0xf)
        prev = 0
        for val in (range(1 << 15) if i % 25 == 24 else range(0, 1 << 15, 37)):   # every offset on four of the hundred CDFs
            sym = L.dvo_cdf_lookup(ctypes.byref(c), val, ctypes.byref(st), ctypes.byref(fr))
            assert prev <= sym
            assert st.value <= val + 1 and val <= st.value + fr.value
            prev = sym
        if i % 25 == 24:
            assert prev == 15


def _simple_rand(state):   # common_tests.rs:46-50
    state = (state * 1103515245 + 12345) & 0xFFFFFFFFFFFFFFFF
    return state, (state // 65536) % 32768


def test_stationary_probability(oracle_blend):
    # common_tests.rs:53-95 (200 000 draws instead of 1 000 000: the model's memory is ~260 blends at its final rate 127 / 32768)
    truth = [(0, 1), (0, 1), (1, 16), (0, 1), (1, 32), (1, 32), (0, 1), (0, 1), (1, 8), (0, 1), (0, 1), (0, 1), (1, 5), (1, 5), (1, 5), (3, 20)]
    cut, acc = [], 0.0
    for a, b in truth:
        acc += np.float32(a) / np.float32(b)
        cut.append(int(round(float(np.float32(CDF_MAX + 1) * np.float32(acc)))))
    assert cut[15] == CDF_MAX + 1
    c = _cdf(oracle_blend)
    seed = 1
    L = oracle_blend.lib()
    for _ in range(200000):
        seed, r = _simple_rand(seed)
        j = next(k for k in range(16) if r < cut[k])
        _blend(oracle_blend, c, j)
        assert 0 <= min(c.c) and max(c.c) <= CDF_MAX                                      # valid(), blend_cdf.rs:172-179
    vals = [L.dvo_cdf_value(ctypes.byref(c), s) for s in range(16)]
    for i, (a, b) in enumerate(truth):
        actual = (vals[i] - (vals[i - 1] if i else 0)) / CDF_MAX
        expected = a / b
        assert abs(expected - actual) < 0.014 or (expected and abs(expected - actual) / expected < 0.15), (i, actual, expected)


def test_nonzero_pdf(oracle_blend):
    # common_tests.rs:98-108 (regression test: symbols that never occur keep a nonzero probability)
    c = _cdf(oracle_blend)
    for _ in range(100000):
        _blend(oracle_blend, c, 15)
    vals = [oracle_blend.lib().dvo_cdf_value(ctypes.byref(c), s) for s in range(16)]
    assert all(vals[i] - (vals[i - 1] if i else 0) > 0 for i in range(15))
    assert c.mix_rate == 127                                                              # the decay stops below 1 << 7 (:203)


def test_mix_rate_schedule(oracle_blend):
    # the kernels read the rate of blend number k from a 512-entry table and count k in sign bits (dv_blend.cuh): the schedule
    # must be stationary well before the counter wraps from 511 to 496
    c = _cdf(oracle_blend)
    rates = []
    for k in range(600):
        rates.append(c.mix_rate)
        _blend(oracle_blend, c, k & 15)
    assert rates[0] == 1536 and rates[386] == 127 and rates[385] == 128 and all(r == 127 for r in rates[386:])


def test_round_trip_and_wire_incompatibility(oracle, oracle_blend):
    from divans_b200 import synth
    text = synth.text_corpus(1 << 17)
    rng = np.random.default_rng(3)
    cases = [b"", b"a", text[:15], text[:4097], text[:70000], rng.integers(0, 256, 5000).astype(np.uint8).tobytes(), bytes(9000)]
    for dcm in (0, 2):
        for raw in cases:
            enc = oracle_blend.encode_raw(raw, oracle_blend.options(dynamic_context_mixing=dcm))
            rc, out = oracle_blend.decode(enc)
            assert rc == 0 and out == raw
    # the default-model decoder cannot read a blend stream (nothing in the stream says which model coded it)
    enc = oracle_blend.encode_raw(text[:4097])
    rc, out = oracle.decode(enc)
    assert rc != 0 or out != text[:4097]
    # LZ77 command streams and random IR through the blend build
    import irfuzz
    for seed in range(6):
        c = oracle_blend.Commands.from_ir(irfuzz.random_ir(oracle_blend, 300 + seed, n_cmds=120, window=16, text=text))
        enc = c.encode(oracle_blend.options(window_size=16, dynamic_context_mixing=seed % 3))
        rc, out = oracle_blend.decode(enc)
        assert rc == 0 and (0, out) == c.recode(16)
