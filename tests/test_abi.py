"""CPU tests of the drop-in boundary: liblvo.so loads, exports every symbol include/lvo.h declares, keeps its struct
layouts in step with the ctypes mirror, and fails loudly (no CPU fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_lvo


def header_functions():
    src = open(os.path.join(ROOT, "include", "lvo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lvo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = load_lvo()
    lib = L.load_library()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lvo.h but not exported by liblvo.so"
    assert sorted(L.EXPORTS) == names, "python binding and header disagree"


def test_header_compiles_as_c_and_struct_sizes_match(tmp_path):
    L = load_lvo()
    prog = tmp_path / "sizes.c"
    prog.write_text('#include "lvo.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lvo_config), sizeof(lvo_stats), '
                    'sizeof(lvo_pose), sizeof(lvo_cloud_view), sizeof(lvo_cloud_out), sizeof(lvo_timings), sizeof(lvo_point));return 0;}\n')
    exe = tmp_path / "sizes"
    assert os.system(f"gcc -std=c99 -Wall -Werror -I{ROOT}/include {prog} -o {exe}") == 0  # plain C, no C++ / torch types in the ABI
    sizes = [int(x) for x in os.popen(str(exe)).read().split()]
    assert sizes == [C.sizeof(L.Config), C.sizeof(L.Stats), C.sizeof(L.Pose), C.sizeof(L.CloudView), C.sizeof(L.CloudOut), C.sizeof(L.Timings), 16]


def test_status_codes_and_defaults():
    L = load_lvo()
    lib = L.load_library()
    cfg = L.Config()
    lib.lvo_default_config(C.byref(cfg))
    # launch/aloam_velodyne_HDL_64.launch:3-13 ; laserOdometry.cpp:364,573 ; HuberLoss(0.1)
    assert (cfg.n_scans, cfg.minimum_range, cfg.line_res, cfg.plane_res, cfg.skip_frame, cfg.outer_iters, cfg.lm_max_iters, cfg.huber, cfg.lanes) == (
        64, 5.0, 0.4, 0.8, 1, 10, 4, 0.1, 1)
    h = C.c_void_p()
    cfg.n_scans = 48   # "wrong scan number", scanRegistration.cpp:203
    assert lib.lvo_create(C.byref(cfg), C.byref(h)) == L.LVO_E_BADARG and not h
    assert lib.lvo_create(None, C.byref(h)) == L.LVO_E_BADARG
    assert lib.lvo_destroy(None) == L.LVO_E_BADARG
    assert lib.lvo_state_bytes() > 1000


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    L = load_lvo()
    with pytest.raises(L.LvoError) as e:
        L.Lvo()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
