"""CPU-side check that the built hot-path objects really are Blackwell tensor-core / TMA code (no GPU needed: `cuobjdump -sass` of the objects
`__graft_entry__.build()` leaves under rust-autograd_b200/csrc/build/).  The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as
proof: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA load / store (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers).
A rebuild that silently fell back to mma.sync / plain loads would keep every parity test green; this one would fail.  profiles/sass_r2.txt is the
committed summary of the same listing (scripts/sass_summary.py)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "rust-autograd_b200", "csrc", "build")

# object -> mnemonics that must occur (the small objects only: the listing of tc_gemm.o / tc_conv.o takes 10+ s each)
EXPECT = {
    "tc_conv_rows.o": ["UTCHMMA", "UTMALDG", "LDTM"],            # window-reuse conv, wide maps (64 -> 64 @128x128 of the benchmark)
    "tc_conv_cols.o": ["UTCHMMA", "UTMALDG", "LDTM"],            # window-reuse conv, narrow maps
    "tc_conv_wgrad_taps.o": ["UTCHMMA", "UTMALDG", "LDTM"],      # all-taps filter gradient
    "tc_conv_first.o": ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"],  # first layer: smem-built im2col tile, TMA-store epilogue
}


def _sass(obj):
    return subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, timeout=120).stdout


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_hot_path_objects_are_tcgen05_and_tma_code(name):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    obj = os.path.join(BUILD, name)
    if not os.path.exists(obj):
        pytest.skip("%s not built here (run __graft_entry__.build())" % name)
    sass = _sass(obj)
    assert "sm_100a" in sass or "SM100" in sass.upper() or "EF_CUDA_SM100" in sass, "not an sm_100a object"
    for m in EXPECT[name]:
        assert sass.count(m) > 0, "%s: no %s instruction in the SASS" % (name, m)
    assert "HMMA.16816" not in sass and "HMMA.1688" not in sass, "%s contains warp-level mma.sync tensor instructions" % name


def test_gemm_object_has_cta_pair_mma():
    """tc_gemm.o: the 256-wide tiles run as cta_group::2 pairs (UTCHMMA.2CTA) with multicast commits, the 3xTF32 tiles as N = 2 TN single-CTA MMAs."""
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    obj = os.path.join(BUILD, "tc_gemm.o")
    if not os.path.exists(obj):
        pytest.skip("tc_gemm.o not built here")
    sass = _sass(obj)
    for m in ["UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMALDG", "LDTM"]:
        assert sass.count(m) > 0, "tc_gemm.o: no %s" % m
    assert sass.count("UTCHMMA") > sass.count("UTCHMMA.2CTA")     # single-CTA kernels (128-wide tiles, 3xTF32) are there too
