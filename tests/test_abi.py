"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and the product path refuses to run without a GPU (no silent CPU fallback)."""
import ctypes
import os

import pytest
import torch

from conftest import PKG, ROOT


def test_header_symbols_exported(native_lib):
    from unidistill_b200 import native

    protos = native.parse_header(os.path.join(ROOT, "include", "unidistill_b200.h"))
    assert len(protos) >= 18
    raw = ctypes.CDLL(native.LIB_PATH)
    missing = [name for name in protos if not hasattr(raw, name)]
    assert not missing, "symbols declared in include/unidistill_b200.h but not exported: %s" % missing


def test_version_and_error_strings(native_lib):
    assert b"sm_100a" in native_lib.ud_version()
    assert native_lib.ud_error_string(0) == b"ok"
    assert b"invalid argument" in native_lib.ud_error_string(1001)
    assert b"workspace" in native_lib.ud_error_string(1002)


def test_library_is_sm100a_only(native_lib):
    """The shipped cubins target sm_100a (no PTX JIT path to other architectures is relied on)."""
    import shutil
    import subprocess

    from unidistill_b200 import native

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_ops_refuse_cpu_tensors(native_lib):
    """Every operator raises on CPU tensors instead of falling back."""
    from unidistill_b200 import sparse, voxelize

    pts = torch.zeros((8, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        voxelize.hard_voxelize([pts], [0.1, 0.1, 0.2], [-1, -1, -1, 1, 1, 1], [20, 20, 10], 10, 100)
    coords = torch.zeros((4, 4), dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        sparse.build_subm_rulebook(coords, 1, [4, 4, 4], 3)


def test_invalid_arguments_are_reported(native_lib):
    """Host-side argument validation happens before any device work (callable without a GPU)."""
    rc = native_lib.ud_sparse_conv_forward(None, None, None, None, 4, 27, 16, 16, None, None)
    assert rc == 1001
    rc = native_lib.ud_bn_apply(None, 4, 16, None, None, None, 1, None, None)
    assert rc == 1001


def test_mirror_modules_import_and_state_dict_keys():
    import numpy as np

    from unidistill.layers.blocks_2d.det3d import BaseBEVBackbone, HeightCompression  # noqa: F401
    from unidistill.layers.blocks_3d.det3d import MeanVFE, VoxelResBackBone8x

    m = VoxelResBackBone8x(5, np.array([1024, 1024, 40]))
    keys = set(m.state_dict().keys())
    # key names and shapes a reference checkpoint carries (spconv-2 weight layout (Cout,kz,ky,kx,Cin))
    assert m.state_dict()["conv_input.0.weight"].shape == (16, 3, 3, 3, 5)
    assert m.state_dict()["conv2.0.0.weight"].shape == (32, 3, 3, 3, 16)
    assert m.state_dict()["conv4.2.conv2.bias"].shape == (128,)
    assert m.state_dict()["conv_out.0.weight"].shape == (128, 3, 1, 1, 128)
    for k in ("conv1.0.bn1.running_mean", "conv3.1.conv1.weight", "conv_out.1.num_batches_tracked"):
        assert k in keys
    assert list(m.sparse_shape) == [41, 1024, 1024]
    assert sum(p.numel() for p in m.parameters()) == 2695312
    assert MeanVFE(5).get_output_feature_dim() == 5
