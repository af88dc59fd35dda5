"""CPU: the C-ABI library loads, exports every symbol the header declares, and validates its
arguments without touching a GPU."""
import ctypes
import os
import re

import pytest

from fast_match_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return ctypes.CDLL(build.LIB)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fastmatch_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert {"fm_top2_u8", "fm_grouped_mutual_u8", "fm_merge_top2", "fm_ratio_f32sqrt", "fm_top2_host_u8",
            "fm_top2_workspace_bytes", "fm_grouped_workspace_bytes", "fm_version", "fm_last_error",
            "fm_device_caps"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), "header declares %s but libfmatch.so does not export it" % s


def test_version_and_argument_validation(lib):
    lib.fm_last_error.restype = ctypes.c_char_p
    assert lib.fm_version() >= 100
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    lib.fm_top2_u8.argtypes = [vp, i64, vp, i64, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_int, vp]
    # negative size / null pointers -> FM_EINVAL (-1) with a message, no CUDA call involved
    assert lib.fm_top2_u8(None, -1, None, 0, 0, None, None, None, None, 0, 0, None) == -1
    assert b"fm_top2_u8" in lib.fm_last_error()
    assert lib.fm_top2_u8(None, 5, None, 5, 0, None, None, None, None, 0, 0, None) == -1
    buf = (ctypes.c_uint8 * 4096)()
    base = ctypes.addressof(buf)
    mis = base + (1 if base % 16 == 0 else 0) + (16 - base % 16) % 16 + 1     # misaligned on purpose
    assert lib.fm_top2_u8(vp(mis), 1, vp(mis), 1, 0, vp(base), vp(base), None, None, 0, 0, None) == -1
    assert b"aligned" in lib.fm_last_error()
    lib.fm_merge_top2.argtypes = [vp, ctypes.c_int32, i64, vp, vp, vp, vp]
    assert lib.fm_merge_top2(None, 2, 10, None, None, None, None) == -1
    lib.fm_grouped_workspace_bytes.restype = ctypes.c_size_t
    lib.fm_grouped_workspace_bytes.argtypes = [i64, i64, i64, ctypes.c_int32]
    assert lib.fm_grouped_workspace_bytes(100, 50, 50, 3) >= 50 * 8
    # M == 0 is a legal no-op
    assert lib.fm_top2_u8(None, 0, None, 0, 0, None, None, None, None, 0, 0, None) == 0


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """The shipped binary must contain the tcgen05 / TMEM / TMA path (not a recompiled legacy one)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build()
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    # tcgen05.mma kind::i8, TMEM loads, TMA loads; and the CTA-pair forms of the dense kernel:
    # cta_group::2 MMA, 2-CTA TMA load, multicast commit
    for mnemonic in ("UTCIMMA", "LDTM", "UTMALDG", "UTCIMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", build.LIB], capture_output=True, text=True).stdout


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/fastmatch_b200.h compiles as strict C99 and a C program that references every entry
    point links against libfmatch.so and runs (argument validation only: no GPU involved)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    build.build()
    names = declared_symbols()
    src = tmp_path / "cabi_check.c"
    src.write_text(
        '#include "fastmatch_b200.h"\n#include <stdio.h>\n#include <string.h>\n'
        "int main(void) {\n"
        "    void *fns[] = {" + ", ".join("(void *)%s" % n for n in names) + "};\n"
        "    int rc = fm_top2_u8(0, -1, 0, 0, 0, 0, 0, 0, 0, 0, FM_ALGO_AUTO, 0);\n"
        '    printf("%d %d %d\\n", fm_version(), (int)(sizeof(fns) / sizeof(fns[0])), rc);\n'
        "    return (rc == FM_EINVAL && strstr(fm_last_error(), \"fm_top2_u8\") != 0) ? 0 : 1;\n}\n")
    exe = tmp_path / "cabi_check"
    libdir = os.path.dirname(build.LIB)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L" + libdir, "-lfmatch", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[1] == str(len(names))
