"""The C-ABI library: loads without a GPU, exports exactly what include/oon_gpu.h declares, fails loudly
when there is no device (no CPU path), and its records have the reference's byte layout.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from out_of_nothing_b200 import capi, snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "oon_gpu.h")).read()


def declared_functions():
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(oon_\w+)\s*\(", HEADER, flags=re.M)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "liboon_gpu.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with include/oon_gpu.h"
    assert L.oon_abi_version() == int(re.search(r"#define OON_ABI_VERSION (\d+)", HEADER).group(1))


def test_no_torch_or_cxx_types_in_the_header():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)              # declarations only, comments stripped
    assert "torch" not in code and "std::" not in code and "at::" not in code and "template" not in code
    assert 'extern "C"' in HEADER


def test_record_layouts_match_the_reference_fixtures():
    assert capi.BODY_F32.itemsize == 60 == snapshot.BODY_F32.itemsize
    assert [capi.BODY_F32.fields[f][1] for f in ("lifetime", "r", "density", "p", "v", "T", "color", "mass", "thrust")] == \
           [4, 8, 12, 16, 24, 32, 36, 40, 44]                                  # SURVEY.md App. A offsets
    assert [snapshot.BODY_F64.fields[f][1] for f in ("lifetime", "r", "density", "p", "v", "T", "color", "mass", "thrust")] == \
           [8, 16, 24, 32, 48, 64, 68, 72, 80]
    assert ctypes.sizeof(capi.Props) == 40 and ctypes.sizeof(capi.Events) == 64


def test_props_default_are_the_reference_defaults():
    L = capi.load()
    p = capi.Props()
    assert L.oon_props_default(ctypes.byref(p)) == capi.OK
    # Properties::defaults, World.cpp:42-50; LoopMode::Default = Half_Matrix, World.hpp:103
    assert abs(p.friction - 0.03) < 1e-9 and p.gravity_mode == capi.GRAVITY_HYPERBOLIC and p.gravity == 6.6743e-11
    assert p.interact_n2n == 0 and p.repulsion_stiffness == 0 and p.collision_mode == capi.COLLISION_GLIDE_THROUGH
    assert p.loop_mode == capi.LOOP_HALF_MATRIX


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_create_fails_loudly_without_a_device():
    L = capi.load()
    h = ctypes.c_void_p()
    rc = L.oon_create(0, ctypes.byref(h))
    assert rc == capi.ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in L.oon_last_error(None)
    from out_of_nothing_b200.world import World
    with pytest.raises(capi.OonError):
        World(0)


def test_null_arguments_are_rejected_not_crashed():
    L = capi.load()
    assert L.oon_props_default(None) == capi.ERR_INVALID
    assert L.oon_destroy(None) == capi.ERR_INVALID
    assert L.oon_step(None, 0.1, 1) == capi.ERR_INVALID
    assert L.oon_get_props(None, None) == capi.ERR_INVALID
    assert L.oon_create(0, None) == capi.ERR_INVALID


def test_shard_range_is_the_partition_the_library_and_the_python_helper_agree_on():
    """oon_shard_range (pure function, no device): the block a rank owns in a freshly uploaded world — the same partition as
    out_of_nothing_b200/sharding.py, tiles the index space, canonical blocks for 1/2/4/8 ranks."""
    import ctypes
    from out_of_nothing_b200 import sharding
    L = capi.load()
    for n in (0, 1, 255, 256, 257, 1010, 10703, 100000, 1000000):
        for nranks in (1, 2, 3, 4, 8):
            at = 0
            for rank in range(nranks):
                first, count = ctypes.c_uint64(), ctypes.c_uint64()
                assert L.oon_shard_range(n, rank, nranks, ctypes.byref(first), ctypes.byref(count)) == capi.OK
                lo, hi = sharding.shard_range(n, rank, nranks)
                assert (first.value, first.value + count.value) == (lo, hi)
                assert first.value == at
                at += count.value
            assert at == n
    assert L.oon_shard_range(10, 2, 2, None, None) == capi.ERR_INVALID
