"""The C-ABI library loads and exports every symbol include/apc.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "apc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from autodriver_pointcloud_preprocessor_b200 import _build, _capi
    _build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 24
    for name in names:
        assert hasattr(lib, name), f"{name} declared in apc.h but not exported"
    assert sorted(_capi.SYMBOLS) == names          # the ctypes binding covers the whole header
    assert lib.apc_version() == 120                 # APC_VERSION in apc.h


def test_struct_layouts_match_header():
    """ctypes mirrors of the header structs: sizes the C compiler agrees with."""
    import subprocess
    import tempfile
    from autodriver_pointcloud_preprocessor_b200 import _capi
    prog = r'''
#include <stdio.h>
#include "apc.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(apc_field), sizeof(apc_cloud_desc), sizeof(apc_filter_cfg),
  sizeof(apc_out_field), sizeof(apc_pipeline_cfg), sizeof(apc_out_mirror), sizeof(apc_pipeline_maps)); return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (_capi.Field, _capi.CloudDesc, _capi.FilterCfg, _capi.OutField, _capi.PipelineCfg,
                                       _capi.OutMirror, _capi.PipelineMaps)]
    assert got == sizes


def test_no_cpu_fallback_without_cuda():
    """Without a GPU the product path refuses to run instead of falling back to the CPU."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from autodriver_pointcloud_preprocessor_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Context(max_points=1024)
