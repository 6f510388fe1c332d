#!/usr/bin/env python3
"""Extract the RFC 7932 static data the divANS path depends on into one blob.

The reference obtains these from the (un-vendored) `brotli ~3.1` crate
(reference: src/codec/dict.rs:4,7 ; src/cmd_to_raw/mod.rs:19-21 ; src/codec/interface.rs:199-238
via src/constants.rs).  They are RFC 7932 constants, so we take them from the system
libbrotlicommon.so.1 (no headers needed) and freeze them in `divans_b200/csrc/brotli_tables.bin`,
which is committed and embedded (.incbin) in both the oracle and the CUDA library.

Blob layout (little endian):
  0      : magic "DVBT" + u32 version(1)
  8      : u32 dict_size (122784), u32 n_transforms (121), u32 prefix_suffix_size, u32 reserved
  24     : u8  size_bits_by_length[32]
  56     : u32 offsets_by_length[32]
  184    : u8  context_lookup[2048]       (mode*512 + {0:lut0,256:lut1}; modes LSB6,MSB6,UTF8,SIGNED)
  2232   : u8  transforms[121*3]          (prefix_id, type, suffix_id) -> padded to 384
  2616   : u16 prefix_suffix_map[50]      -> padded to 128 bytes
  2744   : u8  prefix_suffix[prefix_suffix_size] padded to 256
  3000   : u8  dictionary[dict_size]
"""
import ctypes, struct, sys, os

def main(out_path):
    lib = ctypes.CDLL("libbrotlicommon.so.1")
    class BrotliDictionary(ctypes.Structure):
        _fields_ = [("size_bits_by_length", ctypes.c_uint8 * 32),
                    ("offsets_by_length", ctypes.c_uint32 * 32),
                    ("data_size", ctypes.c_size_t),
                    ("data", ctypes.POINTER(ctypes.c_uint8))]
    class BrotliTransforms(ctypes.Structure):
        _fields_ = [("prefix_suffix_size", ctypes.c_uint16),
                    ("prefix_suffix", ctypes.POINTER(ctypes.c_uint8)),
                    ("prefix_suffix_map", ctypes.POINTER(ctypes.c_uint16)),
                    ("num_transforms", ctypes.c_uint32),
                    ("transforms", ctypes.POINTER(ctypes.c_uint8)),
                    ("params", ctypes.POINTER(ctypes.c_uint8)),
                    ("cutOffTransforms", ctypes.c_int16 * 10)]
    lib.BrotliGetDictionary.restype = ctypes.POINTER(BrotliDictionary)
    lib.BrotliGetTransforms.restype = ctypes.POINTER(BrotliTransforms)
    d = lib.BrotliGetDictionary().contents
    t = lib.BrotliGetTransforms().contents
    assert d.data_size == 122784, d.data_size
    assert t.num_transforms == 121, t.num_transforms
    ctx = (ctypes.c_uint8 * 2048).in_dll(lib, "_kBrotliContextLookupTable")
    dict_bytes = bytes(d.data[i] for i in range(d.data_size))
    tr = bytes(t.transforms[i] for i in range(121 * 3))
    n_ps = max(max(tr[0::3]), max(tr[2::3])) + 1
    assert n_ps <= 64
    psmap = [t.prefix_suffix_map[i] for i in range(n_ps)]
    ps = bytes(t.prefix_suffix[i] for i in range(t.prefix_suffix_size))
    assert len(ps) <= 256
    blob = bytearray()
    blob += b"DVBT" + struct.pack("<I", 1)
    blob += struct.pack("<IIII", d.data_size, 121, len(ps), n_ps)
    blob += bytes(d.size_bits_by_length)
    blob += struct.pack("<32I", *d.offsets_by_length)
    assert len(blob) == 184
    blob += bytes(ctx)
    assert len(blob) == 2232
    blob += tr + bytes(384 - len(tr))
    assert len(blob) == 2616
    pm = struct.pack("<%dH" % n_ps, *psmap)
    blob += pm + bytes(128 - len(pm))
    assert len(blob) == 2744
    blob += ps + bytes(256 - len(ps))
    assert len(blob) == 3000
    blob += dict_bytes
    with open(out_path, "wb") as f:
        f.write(blob)
    print("wrote", out_path, len(blob), "bytes; prefix/suffix entries", n_ps, "ps bytes", len(ps))

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "divans_b200", "csrc", "brotli_tables.bin"))
