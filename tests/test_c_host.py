"""The boundary is a C ABI: include/abcdls.h must be valid C (not only C++), and a plain C host must be able to link
libabcdls.so and call it - the stand-in for the cgo / JNI binding INTEGRATION.md describes.  No GPU needed: only the
calls that do no device work are made (abcdls_create must fail loudly, not fall back, when there is no B200)."""
import os
import shutil
import subprocess

import pytest

from abc_dls_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_HOST = r"""
#include <stdio.h>
#include "abcdls.h"
int main(void) {
    abcdls_handle_t h = 0;
    int rc = abcdls_create(&h, 0);
    printf("version %d\n", abcdls_version());
    printf("create %d %s\n", rc, h ? "handle" : "null");
    printf("nacc %lld\n", (long long)abcdls_abc_single_nacc(100000000, 0.01));
    printf("pitch %d %d\n", abcdls_fused_emit_pitch(16), abcdls_fused_emit_pitch(19));
    printf("tail %d %d\n", abcdls_tail_supported(24, 24), abcdls_tail_supported(40, 3));
    printf("small %zu\n", abcdls_final_small_doubles(16, 16, 1));
    if (h) abcdls_destroy(h);
    return 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_header_is_c_and_a_c_host_links(tmp_path):
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "abcdls.h")],
                   check=True)
    src = tmp_path / "host.c"
    src.write_text(C_HOST)
    exe = tmp_path / "host"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-labcdls",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    kv = {l.split()[0]: l.split()[1:] for l in out}
    assert kv["version"] == ["1"]
    assert kv["nacc"] == ["1000000"] and kv["pitch"] == ["16", "32"] and kv["tail"] == ["1", "0"]
    assert int(kv["small"][0]) > 8 + 16 + 6 * 16
    import torch
    if torch.cuda.is_available():
        assert kv["create"] == ["0", "handle"]
    else:
        assert kv["create"] == [str(_capi.E_NODEVICE), "null"]      # no device: an error code, never a CPU path
