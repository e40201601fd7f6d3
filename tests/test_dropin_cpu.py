"""CPU: the drop-in layer end to end with the CPU oracle standing in for the CUDA engine (tests/_oracle_engine.py):
the SAME frame comparisons against the reference's frames as tests/test_dropin_gpu.py / test_cellchat_gpu.py run on the
B200, executed here through the oracle - so the host logic (input contract, tables, row resolution, scores, consensus,
frame assembly, by_sample, MuData, supplementary columns, unlabeled cells, return_all_lrs) is checked on every
machine, and a difference on the GPU box can only come from the device code."""
import pytest

from liana_py_b200.method import _pipeline as PL
from _oracle_engine import OracleEngine
import test_dropin_gpu as G
import test_cellchat_gpu as GC


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setitem(PL._ENGINES, 0, OracleEngine())
    PL._TABLE_CACHE.clear()
    yield


@pytest.mark.parametrize("name", ["toy700", "small300", "ragged90"])
def test_cellphonedb_and_geometric_mean_frames(name):
    G.test_cellphonedb_frame_identical_to_reference(name)
    G.test_geometric_mean_frame_identical_to_reference(name)


@pytest.mark.parametrize("name", ["small300", "ragged90"])
def test_rank_aggregate_frames(name):
    G.test_rank_aggregate_frame_matches_reference(name)


def test_rank_aggregate_mean_and_noperms():
    G.test_rank_aggregate_mean_and_noperms()


@pytest.mark.parametrize("name", ["pairs", "interactions", "resource", "raw", "layer", "resname"])
def test_call_signature_variants(name):
    G.test_call_signature_variants_identical_to_reference(name)


def test_supp_columns():
    G.test_supp_columns_identical_to_reference()


def test_mudata_input():
    G.test_mudata_input_identical_to_reference()


def test_dense_matrix_input():
    G.test_dense_matrix_input_equals_sparse_input()


def test_non_canonical_matrix():
    G.test_non_canonical_matrix_gives_the_frame_of_its_canonical_form()


def test_cells_without_a_label():
    G.test_cells_without_a_label_identical_to_reference()


def test_signed_values_and_integer_cluster_names():
    G.test_signed_values_and_integer_cluster_names_identical_to_reference()
    G.test_geometric_mean_of_signed_values_follows_the_reference_nan_rules()


def test_return_all_lrs():
    G.test_return_all_lrs_frames_identical_to_reference()
    G.test_return_all_lrs_contract()
    G.test_return_all_lrs_in_the_other_entry_points()


def test_by_sample():
    G.test_by_sample_frames_identical_to_reference()
    G.test_by_sample_matches_per_sample_calls()


def test_consensus_options():
    G.test_consensus_options_identical_to_reference()


def test_consensus_containing_cellchat():
    G.test_consensus_containing_cellchat_reference_quirk()


def test_error_conventions_and_empty_result():
    G.test_error_conventions()
    G.test_nothing_expressed_gives_an_empty_frame()


def test_single_non_permutation_methods():
    G.test_single_non_permutation_methods_run()


@pytest.mark.parametrize("name", ["small300", "ragged90"])
def test_cellchat_and_scseqcomm_frames(name):
    GC.test_cellchat_frame_identical_to_reference(name)
    GC.test_scseqcomm_frame_matches_reference(name)


@pytest.mark.parametrize("method", ["cellchat", "scseqcomm", "rank_aggregate"])
def test_by_sample_of_other_methods(method):
    GC.test_by_sample_of_other_methods_matches_per_sample_calls(method)


def test_tied_subunits():
    G.test_tied_subunits_resolved_like_the_reference()


def test_overlapped_frame_assembly():
    G.test_overlapped_frame_assembly_equals_plain_path()
