"""The C-ABI library loads, exports every symbol include/lama_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lama_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lama_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(api):
    lib = api.lib()
    declared = _declared_symbols()
    assert len(declared) > 50
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(api.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.lama_version()


def test_ctypes_struct_layouts_match_the_header(api, tmp_path):
    import subprocess
    src = tmp_path / "sz.cpp"
    src.write_text('#include "lama_b200.h"\n#include <cstdio>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(lama_device_options), sizeof(lama_pf_options),'
                   ' sizeof(lama_slam_options), sizeof(lama_loc_options));}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["g++", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(api.DeviceOptions), C.sizeof(api.PFOptions), C.sizeof(api.SlamOptions), C.sizeof(api.LocOptions)]


def test_defaults_are_the_reference_defaults(api):
    o = api.PFSlam2D.Options(30)
    # include/lama/pf_slam2d.h:132-185
    assert (o.srr, o.str, o.stt, o.srt) == (0.1, 0.2, 0.1, 0.2)
    assert (o.meas_sigma, o.meas_sigma_gain, o.trans_thresh, o.rot_thresh, o.l2_max) == (0.05, 3.0, 0.5, 0.5, 0.5)
    assert (o.resolution, o.patch_size, o.max_iter, o.truncated_ray, o.truncated_range) == (0.05, 32, 100, 0.0, 0.0)
    s = api.Slam2D.Options()
    assert (s.trans_thresh, s.rot_thresh, s.l2_max, s.resolution, s.patch_size, s.max_iter) == (0.5, 0.5, 0.5, 0.05, 32, 100)
    l = api.Loc2D.Options()
    assert (l.l2_max, l.max_iter, l.resolution) == (1.0, 100, 0.05)   # src/loc2d.cpp:46-58


def test_no_cpu_fallback_without_a_device(api):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.LamaError) as e:
        api.PFSlam2D(api.PFSlam2D.Options(4))
    assert e.value.code == -3
    with pytest.raises(api.LamaError):
        api.Slam2D(api.Slam2D.Options())
    with pytest.raises(api.LamaError):
        api.DynamicDistanceMap()
    with pytest.raises(api.LamaError):
        api.Loc2D(api.Loc2D.Options())


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iris_lama_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "lama_oracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)


def test_header_is_plain_c(tmp_path):
    """include/lama_b200.h is a C header (extern "C" only under __cplusplus): a C99 translation unit compiles, links and runs"""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include "lama_b200.h"\nint main(void){ lama_pf_options o; return lama_pf_options_default(&o) == LAMA_OK && o.particles == 1 ? 0 : 1; }\n')
    exe = tmp_path / "c_abi"
    lib_dir = os.path.join(ROOT, "iris_lama_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir,
                           "-llama_b200", f"-Wl,-rpath,{lib_dir}"])
    assert subprocess.call([str(exe)]) == 0


def test_null_arguments_are_refused_not_dereferenced(api):
    """every getter added in round 2 answers a null handle / null output with LAMA_ERR_ARG and a message (no CUDA call is made on that path)"""
    import ctypes as C
    L = api.lib()
    null = C.c_void_p(None)
    buf = (C.c_uint64 * 4)()
    n = C.c_int(0)
    calls = [
        (L.lama_pf_get_memory_usage, (null, buf)),
        (L.lama_pf_get_timestamps, (null, None, C.c_int(0), C.byref(n))),
        (L.lama_pf_get_summary, (null, None)),
        (L.lama_pf_get_resample_digest, (null, buf)),
        (L.lama_pf_shard_stats, (null, buf)),
        (L.lama_pf_shard_connect, (null, None)),
    ]
    for fn, args in calls:
        fn.restype = C.c_int
        rc = fn(*args)
        assert rc < 0, fn.__name__
        assert len(L.lama_last_error()) > 0
