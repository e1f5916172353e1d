"""CPU-side checks of the C-ABI library: it builds, loads without a GPU, and exports every declared symbol."""
import ctypes
import os

import pytest

from emip_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    syms = _lib.declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_abi_version_matches_header(lib):
    with open(_lib.HEADER_PATH) as f:
        ver = [l for l in f if l.startswith("#define EMIP_ABI_VERSION")][0].split()[-1]
    assert lib.emip_abi_version() == int(ver)


def test_no_runtime_dependency_on_libcuda_or_torch():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libcudart" not in out, out


def test_argument_errors_are_reported_not_thrown(lib):
    lib.emip_last_error.restype = ctypes.c_char_p
    rc = lib.emip_flow_warp_fwd(None, None, None, 1, 3, 8, 8, ctypes.c_longlong(0), ctypes.c_longlong(0), 0, None)
    assert rc == -22 and b"null pointer" in lib.emip_last_error()
    lib.emip_global_matching_workspace.restype = ctypes.c_size_t
    assert lib.emip_global_matching_workspace(16, 128, 44, 44) > 16 * 2 * 1936 * 512


def test_product_path_never_imports_the_oracle():
    pkg = os.path.dirname(build.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from emip_b200.warp import flow_warp
    from emip_b200.matching import global_correlation_softmax
    with pytest.raises(_lib.EmipError):
        flow_warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(_lib.EmipError):
        global_correlation_softmax(torch.zeros(1, 128, 4, 4), torch.zeros(1, 128, 4, 4), True)


def test_injector_state_dict_matches_reference_keys():
    """checkpoints load by key filter, so names and shapes must be the reference's."""
    import cases
    from emip_b200.injector import Injector, PARAM_KEYS
    sd = Injector().state_dict()
    want = {"transformer." + k: shp for k, shp in cases.INJECTOR_SHAPES.items()}
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert list(PARAM_KEYS) == list(cases.INJECTOR_SHAPES)
    assert sum(v.numel() for v in sd.values()) == 206442      # learnable params per injector
