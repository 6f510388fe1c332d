"""Random (valid) IR command streams in the reference's text grammar (src/bin/divans.rs:191-483)."""
import ctypes

import numpy as np


def random_ir(oracle, seed, n_cmds=120, window=16, text=None, wide_speeds=False):
    """wide_speeds: draw the six literal speeds over the whole f8 range a stream can carry (probability/interface.rs:566-585:
    up to 30720), not only the named palette -- i16 counters wrap (frequentist_cdf.rs:74-85) and the LUT divide leaves the
    integer-division regime; such streams are usually not decodable back to their input, but every decoder must walk
    them identically."""
    rng = np.random.default_rng(seed)
    L = oracle.lib()
    lines = ["window %d 0 0 0" % window]
    produced = 0
    ring = 1 << window

    def predmode():
        mode = ["lsb6", "msb6", "utf8", "sign"][int(rng.integers(0, 4))]
        n_bt = int(rng.integers(1, 4))
        n_ctx = int(rng.integers(1, 40))
        lmap = rng.integers(0, n_ctx, 64 * n_bt)
        if rng.random() < 0.2:
            lmap[int(rng.integers(0, lmap.size))] = int(rng.integers(100, 256))   # far value: exercises the 2-nibble escape
        dmap = rng.integers(0, int(rng.integers(1, 9)), 4 * int(rng.integers(1, 4)))
        style = int(rng.integers(0, 4))
        if style == 0:
            mv = np.full(8192, int(rng.integers(0, 9)))
        elif style == 1:
            mv = rng.integers(0, 9, 8192)
        elif style == 2:
            mv = np.repeat(rng.integers(0, 9, 32), 256)
        else:
            mv = np.where(rng.random(8192) < 0.9, 4, rng.integers(0, 9, 8192))
        s = "prediction %s lcontextmap %s dcontextmap %s mixingvalues %s" % (
            mode, " ".join(map(str, lmap)), " ".join(map(str, dmap)), " ".join(map(str, mv)))
        if wide_speeds or rng.random() < 0.5:
            sp = [int(x) for x in rng.choice([1, 2, 4, 8, 16, 32, 48, 64, 128, 512], 6)]
            mx = [int(x) for x in rng.choice([128, 1024, 2048, 4096, 8192, 16384], 6)]
            if wide_speeds:
                f8 = lambda: min(16384, int(L.dvo_u8_to_speed(int(rng.integers(8, 128)))) & 0xffff)   # the IR grammar caps speeds at 16384 (divans.rs:296)
                for k in range(6):
                    if rng.random() < 0.6:
                        sp[k] = f8()
                    if rng.random() < 0.6:
                        mx[k] = f8()
            s += " cmspeedinc %d %d cmspeedmax %d %d stspeedinc %d %d stspeedmax %d %d mxspeedinc %d %d mxspeedmax %d %d" % (
                sp[0], sp[1], mx[0], mx[1], sp[2], sp[3], mx[2], mx[3], sp[4], sp[5], mx[4], mx[5])
        return s, n_bt

    pm, n_bt = predmode()
    lines.append(pm)
    buf = np.zeros(64, np.uint8)
    for _ in range(n_cmds):
        r = rng.random()
        if r < 0.40:
            n = int(rng.choice([1, 2, 3, 7, 14, 15, 16, 17, 30, 200, 1500])) if rng.random() < 0.6 else int(rng.integers(1, 400))
            if text is not None and rng.random() < 0.7:
                o = int(rng.integers(0, len(text) - n))
                data = text[o:o + n]
            else:
                data = rng.integers(0, 256, n).astype(np.uint8).tobytes()
            lines.append("%s %d %s" % ("rndins" if rng.random() < 0.1 else "insert", n, data.hex()))
            produced += n
        elif r < 0.75 and produced > 0:
            maxd = min(produced, ring - 1)
            if rng.random() < 0.5:
                d = int(rng.integers(1, min(maxd, 70) + 1))
            else:
                d = int(rng.integers(1, maxd + 1))
            n = int(rng.choice([1, 2, 3, 4, 5, 9, 14, 15, 16, 31, 100, 700, 5000])) if rng.random() < 0.7 else int(rng.integers(1, 300))
            lines.append("copy %d from %d ctx 0" % (n, d))
            produced += n
        elif r < 0.85:
            ws = int(rng.integers(4, 25))
            bits = [0, 0, 0, 0, 10, 10, 11, 11, 10, 10, 10, 10, 10, 9, 9, 8, 7, 7, 8, 7, 7, 6, 6, 5, 5][ws]
            wid = int(rng.integers(0, 1 << bits))
            t = int(rng.integers(0, 121))
            n = L.dvo_dict_word(ws, wid, t, buf.ctypes.data)
            if n <= 0:
                continue
            lines.append("dict %d word %d,%d 00 func %d 00 ctx 0" % (n, ws, wid, t))
            produced += n
        elif r < 0.90:
            lines.append("ltype %d %d" % (int(rng.integers(0, n_bt)), int(rng.integers(0, 9))))
        elif r < 0.94:
            lines.append("ctype %d" % int(rng.choice([0, 1, 2, 3, 14, 200])))
        elif r < 0.98:
            lines.append("dtype %d" % int(rng.integers(0, 3)))
        else:
            pm, n_bt = predmode()
            lines.append(pm)
    return "\n".join(lines) + "\n"


def random_f8_speeds(oracle, seed):
    """four (inc, lim) pairs over the whole f8 range (up to 30720) for options(literal_adaptation=...): the encoder option that
    overrides the PredictionMode speeds (context_map.rs:144-146)"""
    rng = np.random.default_rng(seed)
    L = oracle.lib()
    return [(int(L.dvo_u8_to_speed(int(rng.integers(8, 128)))), int(L.dvo_u8_to_speed(int(rng.integers(8, 128))))) for _ in range(4)]
