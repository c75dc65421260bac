"""C-ABI checks that need no GPU: the shared library loads, exports every symbol include/splatter360.h
declares, and its pure-host size queries behave."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "splatter360.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s360_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from splatter360_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libsplatter360.so does not export {n}"
    assert set(_lib.EXPORTS) == set(names)


def test_view_struct_layout_and_sizes():
    from splatter360_b200 import _lib
    lib = _lib.load()
    assert ctypes.sizeof(_lib.S360View) == 8 * 4 + 6 * 4 + 4 * 4 + 4 * 8
    assert lib.s360_abi_version() == _lib.ABI_VERSION
    assert lib.s360_geom_bytes(1000) >= 1000 * (48 + 8 + 1)
    assert lib.s360_image_bytes(512, 1024) >= 512 * 1024 * 8 + 2048 * 8
    assert lib.s360_backward_scratch_bytes(10) >= 10 * 9 * 4
    # matrix binning: (chunks of 2048 Gaussians) x tiles counters; beyond 8192 tiles: emit + radix sort (keys, ids, alternates)
    assert lib.s360_binning_scratch_bytes(1 << 20, 1 << 22, 512, 1024) >= 4 * 512 * 2048
    assert lib.s360_binning_scratch_bytes(1 << 20, 1 << 20, 2048, 4096) >= 3 * 4 * (1 << 20)
    assert lib.s360_preprocess_scratch_bytes(1 << 20) >= 4 * 4 * (1 << 20)
    assert b"bad argument" in lib.s360_error_string(-1)
    assert lib.s360_launch_count() >= 0


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    from splatter360_b200 import _lib
    lib = _lib.load()
    v = _lib.S360View()
    v.P, v.image_height, v.image_width, v.mode = 10, 16, 16, 7   # invalid mode, null matrices
    assert lib.s360_mark_visible(ctypes.byref(v), None, None, None) == -1
    assert lib.s360_forward_preprocess(ctypes.byref(v), *([None] * 12)) == -1


def test_batched_entry_points_size_queries_and_argument_checks():
    from splatter360_b200 import _lib
    lib = _lib.load()
    P, cap, V = 1000, 2500, 6
    assert lib.s360_multi_geom_bytes(P, cap) >= cap * (48 + 8 + 1) + 2 * 4 * P + 4
    assert lib.s360_multi_preprocess_scratch_bytes(P, cap) >= 4 * 4 * cap + 4 * (P // 128 + 1)
    assert lib.s360_multi_image_bytes(V, 256, 256) >= V * (256 * 256 * 8 + 256 * 8)
    assert lib.s360_multi_image_bytes(1, 512, 1024) == lib.s360_image_bytes(512, 1024)
    assert lib.s360_multi_binning_scratch_bytes(1 << 20, 1 << 21, 1, 512, 1024) == lib.s360_binning_scratch_bytes(1 << 20, 1 << 21, 512, 1024)
    assert lib.s360_multi_backward_scratch_bytes(cap) >= cap * 9 * 4
    v = _lib.S360View()
    v.P, v.image_height, v.image_width, v.mode, v.scene_scale = 10, 16, 16, 0, 1.0
    # null camera pointers, zero / too many views: rejected before anything is launched
    assert lib.s360_multi_forward_project(ctypes.byref(v), 0, 10, *([None] * 10)) == -1
    assert lib.s360_multi_forward_project(ctypes.byref(v), _lib.MAX_VIEWS + 1, 10, *([None] * 10)) == -1
    assert lib.s360_multi_backward(ctypes.byref(v), 2, 10, *([None] * 9), None, 0, 0.0, 0.0, *([None] * 7)) == -1
    assert lib.s360_cube2equirec_forward(None, None, 0, 1, 3, 8, 16, 32, None, None, None) == -1


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/splatter360.h must be consumable by a C compiler (the boundary is a C-ABI, no C++ or torch types), and a
    C program must link against the library and call its pure-host entry points."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "splatter360.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    from splatter360_b200 import _lib
    _lib.load()
    src = tmp_path / "link.c"
    src.write_text('#include <stdio.h>\n#include "splatter360.h"\n'
                   'int main(void) { printf("%d %zu %s\\n", s360_abi_version(), s360_multi_image_bytes(6, 256, 256), '
                   's360_error_string(S360_ERR_UNSUPPORTED)); return s360_abi_version() == S360_ABI_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "link"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-l:libsplatter360.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == _lib.ABI_VERSION and int(out[1]) > 6 * 256 * 256 * 8


def test_every_kernel_launched_with_programmatic_dependent_launch_waits_first():
    """common.cuh's rule for launch_pdl(): the kernel's FIRST statement is pdl_enter() (griddepcontrol.wait before anything
    touches global memory, executed by every thread) -- otherwise "this grid completed" would no longer imply "everything
    before it completed" for the kernels further down the stream.  Checked on the sources, no GPU needed."""
    import glob
    import re
    src = {p: open(p).read() for p in glob.glob(os.path.join(ROOT, "splatter360_b200", "csrc", "*.cu"))}
    launched = set()
    for text in src.values():
        launched |= set(re.findall(r"launch_pdl\(\s*([A-Za-z_0-9]+)\s*[<,]", text))
    assert len(launched) >= 10, launched
    for name in sorted(launched):
        bodies = []
        for text in src.values():
            for m in re.finditer(r"__global__[^;{]*?\b" + name + r"\s*\(", text, flags=re.S):
                brace = text.index(") {", m.end())
                bodies.append(text[brace + 3:brace + 80].strip())
        assert bodies, f"definition of {name} not found"
        for b in bodies:
            assert b.startswith("pdl_enter();"), f"{name}: first statement is {b[:40]!r}"
