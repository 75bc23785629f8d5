"""The C-ABI product library loads on a CPU-only box and exports every symbol include/dmnd_b200.h declares; it refuses
to create a context without a CUDA device (no CPU fallback).  CPU only, no compute calls."""
import ctypes, os, re
import pytest
from conftest import ROOT, PRODUCT_LIB


def header_functions():
    txt = open(os.path.join(ROOT, "include", "dmnd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dmnd_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from diamond_b200 import api
    assert header_functions() == sorted(api.SYMBOLS)


def test_product_library_exports_every_symbol(product_lib):
    for name in header_functions():
        assert hasattr(product_lib, name), name
    assert product_lib.dmnd_backend().decode() == "cuda-sm100a"


def test_product_library_has_no_cpu_path(product_lib):
    from diamond_b200 import api
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(api.DmndError, match="no CUDA device|CUDA"):
        api.Context(product_lib)


def test_product_library_does_not_contain_the_oracle(product_lib):
    # the oracle identifies itself through dmnd_backend(); the product must not carry its translation unit
    data = open(PRODUCT_LIB, "rb").read()
    assert b"oracle-cpu" not in data


def test_struct_layouts():
    from diamond_b200 import api
    assert ctypes.sizeof(api.Hit) == 16 and api.HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(api.DpProblem) == 16 and api.PROBLEM_DTYPE.itemsize == 16
    assert ctypes.sizeof(api.DpResult) == api.RESULT_DTYPE.itemsize == 56
    assert ctypes.sizeof(api.Match) == api.MATCH_DTYPE.itemsize


def test_ctypes_structs_have_the_c_sizes(tmp_path):
    """diamond_b200/api.py mirrors the structs of include/dmnd_b200.h by hand: a stale mirror corrupts memory silently."""
    import ctypes as C, subprocess
    from conftest import ROOT
    from diamond_b200 import api
    src = tmp_path / "sz.c"
    src.write_text('#include "dmnd_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dmnd_params), sizeof(dmnd_search_opts), '
                   'sizeof(dmnd_run_stats), sizeof(dmnd_match), sizeof(dmnd_hit), sizeof(dmnd_dp_result), sizeof(dmnd_timing));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    got = [C.sizeof(t) for t in (api.Params, api.SearchOpts, api.RunStats, api.Match, api.Hit, api.DpResult, api.Timing)]
    assert got == want


def test_search_opts_layout_matches_the_header(tmp_path):
    """The ctypes mirror of dmnd_search_opts (diamond_b200/api.py) against the C header, field by field: a stale mirror makes the library read
    garbage options (it happened once this round).  gcc prints sizeof and every offsetof; ctypes must agree."""
    import ctypes as C
    import subprocess
    from diamond_b200 import api
    names = [n for n, _ in api.SearchOpts._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dmnd_b200.h"\nint main(void) { printf("%zu\\n", sizeof(dmnd_search_opts));\n'
                   + "".join('printf("%s %%zu\\n", offsetof(dmnd_search_opts, %s));\n' % (n, n) for n in names) + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(out[0]) == C.sizeof(api.SearchOpts)
    for line, n in zip(out[1:], names):
        assert line == f"{n} {getattr(api.SearchOpts, n).offset}"


def test_every_struct_mirror_has_the_header_size(tmp_path):
    """sizeof of every C struct the Python mirror restates (ctypes Structures and numpy record dtypes)."""
    import ctypes as C
    import subprocess
    from diamond_b200 import api
    pairs = {"dmnd_params": C.sizeof(api.Params), "dmnd_hit": C.sizeof(api.Hit), "dmnd_stage_counters": C.sizeof(api.StageCounters), "dmnd_dp_problem": C.sizeof(api.DpProblem),
             "dmnd_dp_result": C.sizeof(api.DpResult), "dmnd_timing": C.sizeof(api.Timing), "dmnd_match": C.sizeof(api.Match), "dmnd_run_stats": C.sizeof(api.RunStats),
             "dmnd_fs_result": api.FS_RESULT_DTYPE.itemsize}
    assert api.MATCH_DTYPE.itemsize == C.sizeof(api.Match)
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "dmnd_b200.h"\nint main(void) {\n' + "".join('printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in pairs) + "return 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert {k: int(v) for k, v in got.items()} == pairs
