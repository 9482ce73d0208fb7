"""CPU checks of the C-ABI boundary: the library loads without a GPU, exports every symbol include/ddm_b200.h
declares (and the ctypes table matches it), refuses to run without a B200, and the product never imports oracle/."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import diffusion_models_b200 as ddm
    ddm._lib.build()
    return ddm._lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "ddm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    import diffusion_models_b200 as ddm
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ddm_b200.h but not exported"
    assert sorted(ddm._lib.EXPORTS) == names, "ctypes table and header disagree"


def test_abi_version_and_error_strings(lib):
    assert lib.ddm_abi_version() == 2
    assert b"ddm_init" in lib.ddm_error_string(-1)
    assert lib.ddm_error_string(-3).decode().startswith("unsupported")


def test_conv_args_struct_layout_matches_header():
    """sizeof(ddm_conv_args) as compiled by gcc from the header == ctypes mirror."""
    import ctypes
    import diffusion_models_b200 as ddm
    prog = '#include <stdio.h>\n#include "ddm_b200.h"\nint main(){printf("%zu", sizeof(ddm_conv_args));return 0;}'
    exe = os.path.join(ROOT, "tests", "_sizeof_test")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    try:
        size = int(subprocess.run([exe], capture_output=True, check=True).stdout)
    finally:
        os.remove(exe)
    assert size == ctypes.sizeof(ddm._lib.ConvArgs)


def test_no_gpu_no_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import diffusion_models_b200 as ddm
    assert lib.ddm_init(0) != 0                                   # fails loudly
    assert lib.ddm_randn(None, 0, 0, 16, None) == -1              # every launch refuses before ddm_init succeeds
    m = ddm.Unet(dim=32, dim_mults=(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))
    d = ddm.DenoisingDiffusion(m, image_size=16, sampling_timesteps=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.sample(batch_size=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffusion-models_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "fake_lib" not in text and "kernel_ref" not in text, f


def test_flop_model_matches_survey():
    from diffusion_models_b200.arch import build_spec
    from diffusion_models_b200.flops import unet_flops_per_image as f
    s = build_spec(64)
    assert abs(f(s, 32, 32) / 1e9 - 3.606) < 1e-3 and abs(f(s, 64, 64) / 1e9 - 14.411) < 1e-3
    assert abs(f(build_spec(64, channels=4, cond_channels=4), 64, 64) / 1e9 - 14.540) < 1e-3
    assert abs(f(build_spec(64, channels=4, text_mode="xattn"), 64, 64, 77) / 1e9 - 14.548) < 1e-3
