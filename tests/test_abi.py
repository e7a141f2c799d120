"""The C-ABI library loads and exports every symbol include/imagew_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from imageworsener_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "imagew_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(iw_[a-z0-9_]+|iwb200_[a-z0-9_]+)\s*\(", text)
    skip = {"iw_mallocfn_type", "iw_freefn_type", "iw_warningfn_type", "iw_translatefn_type"}
    return sorted(set(n for n in names if n not in skip))


def test_product_library_exports_every_declared_symbol():
    so = _build.build_product()
    dll = ctypes.CDLL(so, mode=os.RTLD_LOCAL)
    missing = [n for n in declared_functions() if not hasattr(dll, n)]
    assert not missing, missing
    assert len(declared_functions()) > 80


def test_version_string_names_the_target():
    dll = ctypes.CDLL(_build.build_product(), mode=os.RTLD_LOCAL)
    dll.iwb200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.iwb200_version()


def test_product_contains_no_cpu_path():
    """The shipped library must not contain the emulation layer, and must refuse other backends."""
    dll = ctypes.CDLL(_build.build_product(), mode=os.RTLD_LOCAL)
    dll.iwb200_version.restype = ctypes.c_char_p
    assert b"emulation" not in dll.iwb200_version()
    src = open(os.path.join(ROOT, "imageworsener_b200", "_build.py")).read()
    assert "-DIWB_EMU" not in src.split("def build_emulator")[0]


def test_single_use_and_error_convention(emu):
    lib = emu
    ctx = lib.iw_create_context(None)
    assert lib.iw_process_image(ctx) == 0          # no output profile set
    assert lib.iw_get_errorflag(ctx) == 1
    assert "Output profile not set" in lib.errmsg(ctx)
    assert lib.iw_process_image(ctx) == 0          # contexts are single use; first error sticks
    assert "Output profile not set" in lib.errmsg(ctx)
    lib.iw_destroy_context(ctx)


def test_option_dictionary(emu):
    lib = emu
    ctx = lib.iw_create_context(None)
    lib.iw_set_option(ctx, b"backend", b"cuda")
    lib.iw_set_option(ctx, b"flag", b"")
    assert lib.iw_get_option(ctx, b"backend") == b"cuda"
    assert lib.iw_get_option(ctx, b"flag") == b"1"
    assert lib.iw_get_option(ctx, b"nope") is None
    lib.iw_set_option(ctx, b"backend", b"cpu")
    assert lib.iw_get_option(ctx, b"backend") == b"cpu"
    lib.iw_destroy_context(ctx)


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """include/imagew_b200.h is the boundary for C callers (the reference is C): it must compile as C99 with
    warnings on, and a plain C program using the public calls (tools/batch_driver.c) must link against the library."""
    import subprocess
    so = _build.build_product()
    hdr_only = tmp_path / "hdr.c"
    hdr_only.write_text('#include "imagew_b200.h"\nint main(void) { return IW_VERSION_INT == 0x010305 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-c", str(hdr_only), "-o", str(tmp_path / "hdr.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = tmp_path / "batch_driver"
    r = subprocess.run(["gcc", "-O1", "-pthread", "-Wall", "-I", inc, os.path.join(ROOT, "tools", "batch_driver.c"), "-o", str(exe),
                        "-L", os.path.dirname(so), "-l:" + os.path.basename(so), "-Wl,-rpath," + os.path.dirname(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
