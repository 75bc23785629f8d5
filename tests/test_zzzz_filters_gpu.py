"""Report filters (--id / --query-cover / --subject-cover) on the device: the files of tests/test_filters.py through the product CLI.
The filtered schedule is host logic over the same kernels (score-only and traceback DP in more, smaller waves; the seed stage's hits
thinned by the length ratio for equal covers), so the reference's goldens must come out byte for byte here as well.
(Written after the round's GPU budget was spent: first device run = the round-end suite.)"""
import os
import pytest
from conftest import GOLDEN, ROOT
from test_filters import PROTEIN, TRANSLATED, run_protein, run_translated

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "diamond_b200", "bin", "dmnd-b200")


@pytest.mark.parametrize("name,lvl,flags", PROTEIN)
def test_filtered_protein_search_on_device(product_lib, name, lvl, flags, tmp_path):
    assert run_protein(CLI, name, flags, tmp_path) == open(os.path.join(GOLDEN, f"{name}.{lvl}.tsv")).read()


@pytest.mark.parametrize("lvl,flags", TRANSLATED)
def test_filtered_translated_search_on_device(product_lib, lvl, flags, tmp_path):
    assert run_translated(CLI, flags, tmp_path) == open(os.path.join(GOLDEN, f"bx.{lvl}.tsv")).read()


# ---- BLAST XML and the blocked run's unaligned records through the product CLI (tests/test_formats.py holds the CPU twins)
def test_xml_and_blocked_unaligned_on_device(product_lib, tmp_path):
    from test_formats import check_blocked_unaligned, check_daa, check_xml_protein, check_xml_translated
    check_xml_protein(CLI, tmp_path)
    check_xml_translated(CLI, tmp_path)
    check_blocked_unaligned(CLI, tmp_path)
    check_daa(CLI, tmp_path)


def test_frameshift_formats_on_device(product_lib, tmp_path):
    from test_formats import FRAMESHIFT_FORMATS, check_frameshift_format
    for lvl, ext, flags in FRAMESHIFT_FORMATS:
        check_frameshift_format(CLI, lvl, ext, flags, tmp_path)


def test_no_self_hits_on_device(product_lib, tmp_path):
    from test_filters import check_no_self_hits
    check_no_self_hits(CLI, tmp_path)


def test_view_on_device_build(product_lib, tmp_path):
    """`view` needs no GPU, but the product CLI and library must carry it (dmnd_alignment_stats) like the test stand-in does."""
    from test_formats import check_view
    check_view(CLI, tmp_path)
