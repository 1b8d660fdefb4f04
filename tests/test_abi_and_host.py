"""CPU-side checks: the C-ABI library loads and exports every symbol include/swgpu.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side helpers mirror the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "swgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import native
    lib = sw.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    bound = {n for n, _, _ in native.SYMBOLS}
    for name in names:
        assert hasattr(lib, name), "libswgpu.so does not export " + name
        assert name in bound, "python binding is missing " + name


def test_no_cpu_fallback_without_gpu():
    import torch
    import schwarzwald_b200 as sw
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sw.SwgpuError) as e:
        sw.GpuTiler("RANDOM_GRID", "FAST", [0, 0, 0], [1, 1, 1], 0.01)
    assert e.value.code == 2  # SW_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    """Nothing under schwarzwald_b200/ may import, link or execute oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "schwarzwald_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sworacle" not in text and "oracle/" not in text and "swo_" not in text, f


def test_struct_layouts_match_header():
    from oracle import sworacle
    from schwarzwald_b200 import native, tiler
    assert ctypes.sizeof(native.SwParams) == ctypes.sizeof(sworacle.SwParams) == 80
    assert tiler.NODE_DTYPE.itemsize == sworacle.NODE_DTYPE.itemsize == 32
    assert ctypes.sizeof(native.SwgpuStats) == 144


def test_cubic_bounds_and_spacing_follow_reference():
    """AABB::makeCubic (math/AABB.h:50-61) and spacing = (float)(diagonal / 250)
    (process/TilerProcess.cpp:598-604)."""
    import schwarzwald_b200 as sw
    mn, mx = sw.cubic_bounds([0.0, 10.0, -5.0], [100.0, 40.0, 5.0])
    assert mn.tolist() == [0.0, -25.0, -50.0] and mx.tolist() == [100.0, 75.0, 50.0]
    s = sw.spacing_from_diagonal_fraction(mn, mx)
    assert s.dtype == np.float32 and s == np.float32(np.sqrt(3 * 100.0 ** 2) / 250)
    omn, omx = sw.cubic_bounds_at_origin([0.0, 10.0, -5.0], [100.0, 40.0, 5.0])
    assert omn.tolist() == [-50.0] * 3 and omx.tolist() == [50.0] * 3


def test_synthetic_generators_are_deterministic_and_chunk_invariant():
    import torch
    from schwarzwald_b200 import synth
    for kind, seed in (("uniform", 1), ("terrain", 2), ("urban", 3), ("skewed", 5)):
        a = synth.generate(kind, 30_000, seed, chunk=7_000)
        b = synth.generate(kind, 30_000, seed, chunk=1 << 20)
        assert torch.equal(a, b)
        assert a.dtype == torch.float64 and a.shape == (30_000, 3)
        # LAS-like: coordinates are integer millimetres times the scale plus an offset
    x = synth.generate("skewed", 200_000, 5, side_m=100.0)
    inside = ((x[:, 0] >= 31.0) & (x[:, 0] < 31.0 + 21.54)).float().mean().item()
    assert inside > 0.94


def test_header_is_plain_c_and_struct_sizes_match_the_bindings(tmp_path):
    """include/swgpu.h must be consumable from C (the boundary is a C ABI) and the ctypes / numpy mirrors must
    have the sizes the C compiler gives the structs."""
    import subprocess
    from oracle import sworacle
    from schwarzwald_b200 import native, tiler
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "swgpu.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(sw_params), sizeof(sw_node), sizeof(sw_las_transform),\n'
        '                        sizeof(sw_las_node_header), sizeof(swgpu_stats)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(native.SwParams), tiler.NODE_DTYPE.itemsize, ctypes.sizeof(native.SwLasTransform),
                     tiler.LAS_HEADER_DTYPE.itemsize, ctypes.sizeof(native.SwgpuStats)]
    assert ctypes.sizeof(sworacle.SwLasTransform) == sizes[2] and sworacle.LAS_HEADER_DTYPE.itemsize == sizes[3]


def test_cpp_host_layer_compiles_against_the_header(tmp_path):
    """schwarzwald_b200/host/swgpu_tiler.hpp (the RAII layer the reference-side adapter builds on) instantiated
    with every method, syntax and types only."""
    import subprocess
    src = tmp_path / "host.cpp"
    src.write_text(
        '#include "swgpu_tiler.hpp"\n'
        'void use(swgpu::Tiler& t, double* xyz, const int32_t* las, const sw_las_transform& tr) {\n'
        '  t.index_batch(xyz, 3); t.index_batch_las(las, 3, tr); t.positions(xyz); t.finalize();\n'
        '  auto r = t.result(); auto p = t.payload_pnts(); auto l = t.payload_las(); (void)t.start_level();\n'
        '  (void)r; (void)p; (void)l; (void)swgpu::node_name(5, 1); (void)swgpu::sampling_from_name("JITTERED"); }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                    "-I", os.path.join(ROOT, "schwarzwald_b200", "host"), str(src)], check=True)
