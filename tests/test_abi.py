"""The C-ABI library loads here (no GPU) and exports every symbol include/cnhead.h declares; the
ctypes mirrors of the argument structs have the layout the header describes."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import PKG, ROOT

HEADER = os.path.join(ROOT, "include", "cnhead.h")


@pytest.fixture(scope="module")
def library():
    from cnhead import _lib as L
    if not os.path.exists(L.LIB_PATH):
        subprocess.run(["bash", os.path.join(PKG, "csrc", "build.sh")], check=True)
    return L.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cnh_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("cnh_detloss_fused", "cnh_detloss_count", "cnh_detloss_main", "cnh_detloss_finalize",
                 "cnh_scale_inplace", "cnh_softmax_loss", "cnh_entropy_map_fwd", "cnh_entropy_map_bwd",
                 "cnh_bce_const", "cnh_decode", "cnh_decode_workspace_bytes", "cnh_raster_targets", "cnh_version",
                 "cnh_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(library):
    for name in declared_symbols():
        assert hasattr(library, name), name
    assert library.cnh_version() == 103


def test_struct_layouts_match_header():
    from cnhead import _lib as L
    # sizes follow from the C declarations (LP64): checked against a C compiler
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "cnhead.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %d\n", sizeof(cnh_head), sizeof(cnh_detloss_args), sizeof(cnh_scale_args),
             sizeof(cnh_decode_args), offsetof(cnh_detloss_args, heads), offsetof(cnh_decode_args, apply_sigmoid),
             offsetof(cnh_decode_args, counts_out), sizeof(cnh_raster_args), offsetof(cnh_raster_args, boxes),
             sizeof(cnh_peers), offsetof(cnh_peers, status), CNH_MAILBOX_BYTES);
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    got = [C.sizeof(L.Head), C.sizeof(L.DetLossArgs), C.sizeof(L.ScaleArgs), C.sizeof(L.DecodeArgs),
           L.DetLossArgs.heads.offset, L.DecodeArgs.apply_sigmoid.offset, L.DecodeArgs.counts_out.offset,
           C.sizeof(L.RasterArgs), L.RasterArgs.boxes.offset, C.sizeof(L.Peers), L.Peers.status.offset,
           L.MAILBOX_BYTES]
    assert [int(v) for v in out] == got


def test_argument_errors_are_reported_without_a_gpu(library):
    from cnhead import _lib as L
    a = L.DetLossArgs()
    assert library.cnh_detloss_workspace_bytes(C.byref(a)) == 0          # B == 0
    assert b"bad dims" in library.cnh_last_error()
    d = L.DecodeArgs()
    d.B, d.C, d.H, d.W, d.K, d.D = 1, 1, 4, 4, 17, 2
    assert library.cnh_decode_workspace_bytes(C.byref(d)) == 0
    assert b"K=17" in library.cnh_last_error()
    assert library.cnh_bce_const(None, None, None, 4, 1.0, None) == -1
    r = L.RasterArgs()
    assert library.cnh_raster_targets(C.byref(r), None) == -2 and b"raster_targets" in library.cnh_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from cnhead import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libcnhead_sm100.so")
    with pytest.raises(RuntimeError, match="not built"):
        L.lib()
