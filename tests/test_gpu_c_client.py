"""-m gpu: a C program (tests/c_client/ffi_client.c, compiled here with gcc) drives the reference's FFI surface of
libdivans_b200.so the way the reference's own c/example.c does: custom CAllocator (checking heap wrapper and LIFO bump
arena), the 2-parameter constructor declaration, divans_set_option, 1/15/65536-byte buffers, round trip."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    import divans_b200
    divans_b200.load_library()
    libdir = os.path.dirname(divans_b200.LIB_PATH)
    exe = str(tmp_path_factory.mktemp("c_client") / "ffi_client")
    subprocess.check_call(["gcc", "-O1", "-g", "-Wall", "-o", exe, os.path.join(HERE, "c_client", "ffi_client.c"),
                           "-L" + libdir, "-ldivans_b200", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.parametrize("mode", ["heap", "arena"])
def test_c_client_round_trip_through_custom_allocator(client, mode):
    p = subprocess.run([client, mode], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all ok" in p.stdout
