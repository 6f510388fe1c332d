"""Deterministic synthetic inputs for tests and bench (no dataset access on the GPU box).

`text_streams` makes English-like text: a Zipf-distributed vocabulary of pseudo-words built from English letter
frequencies, sentences with punctuation/capitals.  Under the divANS literal model it compresses to ~0.4-0.5 like
the Canterbury text files the reference tests with (src/bin/integration_test.rs:232-236).
`bernoulli_streams` makes the BASELINE.json config-5 inputs: every BIT of every byte is 1 with probability 1-p.
"""
import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5, 1.0, 0.8,
                      0.15, 0.15, 0.10, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


def _vocab(rng, n_words=4096):
    lens = np.clip(rng.poisson(4.2, n_words) + 1, 1, 14)
    words = []
    for L in lens:
        words.append(bytes(rng.choice(_LETTERS, size=int(L), p=_LETTER_P)))
    return words


def text_corpus(n_bytes, seed=0xD1FA15):
    rng = np.random.default_rng(seed)
    words = _vocab(rng)
    ranks = np.arange(1, len(words) + 1, dtype=np.float64)
    p = 1.0 / ranks ** 1.05
    p /= p.sum()
    out = bytearray()
    sent_left = 0
    cap = True
    while len(out) < n_bytes:
        idx = rng.choice(len(words), size=4096, p=p)
        u = rng.random(4096)
        for k, w in enumerate(idx):
            wd = words[w]
            if cap:
                wd = wd[:1].upper() + wd[1:]
                cap = False
            out += wd
            if sent_left <= 0:
                out += b". " if u[k] < 0.8 else (b"? " if u[k] < 0.9 else b"!\n")
                sent_left = 4 + int(u[k] * 97) % 17
                cap = True
            else:
                out += b", " if u[k] < 0.07 else b" "
                sent_left -= 1
    return bytes(out[:n_bytes])


def text_streams(n_streams, stream_bytes, seed=0xD1FA15, corpus_bytes=1 << 22):
    """(blob uint8 [n*stream_bytes], offsets, lengths): stream i is a window of the corpus at a pseudo-random offset with
    ~1.5% of positions replaced by random printable bytes so that streams are not substrings of each other."""
    corpus = np.frombuffer(text_corpus(max(corpus_bytes, stream_bytes * 2), seed), dtype=np.uint8)
    rng = np.random.default_rng(seed ^ 0x5EED)
    starts = (np.arange(n_streams, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(corpus.size - stream_bytes)
    blob = np.empty(n_streams * stream_bytes, np.uint8)
    for i in range(n_streams):
        s = int(starts[i])
        blob[i * stream_bytes:(i + 1) * stream_bytes] = corpus[s:s + stream_bytes]
    n_mut = int(blob.size * (1.0 / 64))
    pos = rng.integers(0, blob.size, n_mut)
    blob[pos] = rng.integers(32, 127, n_mut).astype(np.uint8)
    off = np.arange(n_streams, dtype=np.uint64) * np.uint64(stream_bytes)
    ln = np.full(n_streams, stream_bytes, np.uint64)
    return blob, off, ln


def bernoulli_streams(n_streams, stream_bytes, p, seed=0xB17):
    rng = np.random.default_rng(seed)
    bits = (rng.random((n_streams * stream_bytes, 8)) >= p).astype(np.uint8)
    blob = np.packbits(bits, axis=1, bitorder="little").reshape(-1)
    off = np.arange(n_streams, dtype=np.uint64) * np.uint64(stream_bytes)
    ln = np.full(n_streams, stream_bytes, np.uint64)
    return blob, off, ln
