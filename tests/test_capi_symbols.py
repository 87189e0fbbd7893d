"""The C-ABI library loads on a CPU-only machine and exports every symbol that
include/fidget_cuda.h and csrc/host/host_capi.h declare.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_library_exports_every_declared_symbol():
    from fidget_b200 import _lib
    lib = _lib.load()
    names = declared("include/fidget_cuda.h", "fc_") + declared("fidget_b200/csrc/host/host_capi.h", "fh_")
    assert len(names) > 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the Python binding table covers the whole CUDA header too
    assert sorted(_lib.CUDA_API) == declared("include/fidget_cuda.h", "fc_")
    assert lib.fc_abi_version() == 2


def test_struct_layouts_match_header():
    from fidget_b200 import _lib
    assert C.sizeof(_lib.FcTapeInfo) == 7 * 4
    assert C.sizeof(_lib.FcRender2dCfg) == 4 * 2 + 64 + 4 + 4 + 4 + 32 + 4 + 8 + 4 + 64 + 8 + 4
    assert C.sizeof(_lib.FcRender3dCfg) == 4 * 3 + 64 + 4 + 32 + 4 + 8 + 4 + 64 + 8 + 8
    assert C.sizeof(_lib.FcRenderStats) == 5 * 64 + 3 * 8 + 4 + 16 * 4 + 4  # + tail padding to 8


def test_no_silent_cpu_fallback():
    """Without a usable GPU the backend must fail loudly (FC_ERR_NO_DEVICE)."""
    import fidget_b200 as fb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(fb.CudaError) as e:
        fb.CudaContext(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under fidget_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fidget_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"(?m)^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", text):
                    bad.append(f)
    assert not bad, bad


def test_header_is_plain_c_and_layouts_agree(tmp_path):
    """include/fidget_cuda.h must be consumable by a C compiler (it is what cgo / bindgen / JNI read), and the
    struct sizes the C compiler sees must be the ones the ctypes mirror declares."""
    import subprocess
    from fidget_b200 import _lib
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "fidget_cuda.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fc_tape_info), sizeof(fc_render2d_cfg), sizeof(fc_render3d_cfg),\n'
                   '         sizeof(fc_render_stats), sizeof(fc_octree_leaf), sizeof(fc_octree_cfg), sizeof(fc_schedule_info));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "hdr"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", str(exe), str(src)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    mirror = [C.sizeof(t) for t in (_lib.FcTapeInfo, _lib.FcRender2dCfg, _lib.FcRender3dCfg, _lib.FcRenderStats,
                                    None, _lib.FcOctreeCfg, _lib.FcScheduleInfo) if t is not None]
    assert sizes[:4] == mirror[:4] and sizes[5:] == mirror[4:]
    assert sizes[4] == 348


def build_c_example(tmp_path):
    import subprocess
    exe = tmp_path / "render2d"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "fidget_b200", "csrc", "host"), os.path.join(ROOT, "examples", "render2d.c"),
                           "-L", os.path.join(ROOT, "fidget_b200"), "-lfidget_cuda",
                           "-Wl,-rpath," + os.path.join(ROOT, "fidget_b200"), "-o", str(exe)])
    return exe


def test_c_client_links_against_the_abi(tmp_path):
    """examples/render2d.c uses nothing but the two C headers; it must compile as C99 and link."""
    assert build_c_example(tmp_path).exists()


def test_generated_rust_binding_is_current():
    """bindings/rust/ffi.rs is generated from the header; it must not be stale and must declare every fc_* symbol."""
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "--check"]) == 0
    rs = open(os.path.join(ROOT, "bindings", "rust", "ffi.rs")).read()
    for name in declared("include/fidget_cuda.h", "fc_"):
        assert f"pub fn {name}(" in rs, name


def test_python_constants_mirror_the_header_defines():
    """Every FC_FLAG_* / FC_OUT_* / FC_ERR_* / FC_ABI_VERSION #define of include/fidget_cuda.h that the ctypes face
    names carries the header's value (a flag added on one side only would silently select another behaviour)."""
    import re
    from fidget_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "include", "fidget_cuda.h")) as f:
        text = f.read()
    defines = {m.group(1): int(m.group(2).rstrip("uU"), 0)
               for m in re.finditer(r"^#define\s+(FC_[A-Z0-9_]+)\s+(-?(?:0x[0-9a-fA-F]+|\d+)[uU]?)\b", text, re.M)}
    flags = {k: v for k, v in defines.items() if k.startswith("FC_FLAG_")}
    assert len(flags) >= 6 and len(set(flags.values())) == len(flags)           # distinct bits
    assert all(v & (v - 1) == 0 for v in flags.values())
    checked = 0
    for name, value in defines.items():
        if hasattr(_lib, name):
            assert getattr(_lib, name) == value, name
            checked += 1
    assert checked >= 10
    for name in flags:                                                            # every flag is reachable from Python
        assert hasattr(_lib, name), name
