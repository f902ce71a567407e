"""The C-ABI library builds for sm_100a without a GPU, loads, and exports exactly what include/s2s_b200.h declares
(no compute calls here: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "s2s_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2s_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_bound_symbols():
    from stain2stain_b200 import _lib
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_builds_loads_and_exports_every_symbol():
    from stain2stain_b200 import _build, _lib
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in include/s2s_b200.h but not exported"
    loaded = _lib.load(build_if_missing=False)
    assert loaded.s2s_abi_version() == 1
    assert loaded.s2s_adam_chunk() > 0
    assert ctypes.sizeof(_lib.ConvSrc) == 24  # {void*, int, int, int} + padding, as in the header


def test_sass_is_blackwell_native():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG (B200_PROFILING.md evidence table)."""
    from stain2stain_b200 import _build
    path = _build.build()
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    # the register-level mma.sync (HMMA) instruction is allowed in exactly two places: the flash-style attention core
    # (csrc/attention.cuh explains why) and the 3-output-channel head conv (csrc/head_conv.cuh: N = 3, memory-bound); every
    # other conv / GEMM of the path must be tcgen05
    for sec in sass.split("Function : ")[1:]:
        name = sec.split("\n", 1)[0]
        if "HMMA." in sec.replace("UTCHMMA", ""):
            assert "attn_" in name or "head_conv" in name, f"legacy mma.sync found outside attention / head conv: {name}"


def test_no_cpu_fallback():
    import torch
    from stain2stain_b200 import _lib, kernels
    with pytest.raises(_lib.S2SError):
        _lib.ptr(torch.zeros(4))
    from stain2stain_b200.unet import UNetModel
    net = UNetModel(dim=[3, 32, 32], num_channels=32, num_res_blocks=1, attention_resolutions="16",
                    use_scale_shift_norm=True, num_head_channels=16, channel_mult=[1, 2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1), torch.zeros(1, 3, 32, 32))
    assert kernels.ACT in (kernels.FMT_BF16, kernels.FMT_F16)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stain2stain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_host_side_geometry_queries_need_no_gpu():
    """The work-decomposition queries of the C ABI are pure host arithmetic (no device needed): the chunking of the streaming
    norm kernels (a function of (B, HW) only, never fewer than 128 / 256 pixels per CTA, covers the sample), the tile count of
    the weight packing and the slice count of the two-level BatchNorm folds."""
    from stain2stain_b200 import _lib
    L = _lib.load()
    for B, HW in [(64, 256 * 256), (64, 128 * 128), (64, 64 * 64), (64, 32 * 32), (2, 256 * 256), (16, 512 * 512), (1, 16 * 16),
                  (3, 24 * 40)]:
        chunks = L.s2s_gn_chunks(B, HW)
        assert chunks >= 1
        ppc = -(-HW // chunks)
        assert chunks * ppc >= HW                                  # the chunks cover the sample
        assert ppc >= min(HW, 256 if HW >= 4096 else 128) or chunks == 1, (B, HW, chunks)
        assert L.s2s_gn_chunks(B, HW) == chunks                    # deterministic: concat sources are cut alike
    # weight packing: tiles of 16 destination rows x 64 destination-contiguous elements
    assert L.s2s_pack_tiles(128, 128, 0) == 8 * 2 and L.s2s_pack_tiles(128, 128, 1) == 8 * 2
    assert L.s2s_pack_tiles(96, 40, 0) == 6 * 1 and L.s2s_pack_tiles(96, 40, 1) == 3 * 2
    assert L.s2s_pack_tiles(0, 40, 0) == 0
    # two-level folds: at least 256 partials per slice, at most 128 slices
    assert [L.s2s_bn_fold_slices(n) for n in (1, 255, 256, 1776, 32768, 10 ** 6)] == [1, 1, 1, 6, 128, 128]


def test_header_is_plain_c_and_a_c_program_binds_the_library(tmp_path):
    """The boundary is a C ABI, not a C++ or torch one: include/s2s_b200.h compiles as C99 and a C program linked against
    the library calls its host-only entry points (what a cgo / JNI / N-API binding of another host language would do)."""
    from stain2stain_b200 import _build
    lib = _build.build()
    src = tmp_path / "bind.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "s2s_b200.h"
int main(void) {
    s2s_conv_src s;  /* the structs of the header are plain C aggregates */
    memset(&s, 0, sizeof s);
    printf("%d %d %d %d %d %d %zu\n", s2s_abi_version(), s2s_adam_chunk() > 0, s2s_linear_max_jobs() > 0,
           s2s_attn_supported(32), s2s_attn_supported(48), s2s_gn_chunks(64, 256 * 256), sizeof s);
    return s2s_last_error() == NULL;
}
''')
    exe = tmp_path / "bind"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-comment", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe),
                    "-L", os.path.dirname(lib), "-ls2s_b200", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[:5] == ["1", "1", "1", "1", "0"], out
    assert int(out[5]) >= 1 and int(out[6]) == 24
