"""Pins the search/filter oracle against every known-answer test the reference
holds for half B of the hot path.  CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import search_oracle as so

GOLDEN = Path(__file__).parent / "golden"

ROWS_STORAGE = [
    # reference tests/test_storage.py:107-151
    dict(id="chunk_001", session_id="session_1", project_name="test_project", chunk_type="qa_pair",
         timestamp="2024-01-15T10:00:00", has_code=0, has_tools=0, message_count=2, char_count=38, word_count=8),
    dict(id="chunk_002", session_id="session_1", project_name="test_project", chunk_type="code_block",
         timestamp="2024-01-15T10:01:00", has_code=1, has_tools=1, message_count=1, char_count=43, word_count=6),
    dict(id="chunk_003", session_id="session_2", project_name="other_project", chunk_type="tool_usage",
         timestamp="2024-01-15T11:00:00", has_code=0, has_tools=1, message_count=3, char_count=37, word_count=4),
]
X_STORAGE = np.array([[.1, .2, .3, .4], [.5, .6, .7, .8], [.9, .1, .2, .3]], np.float32)


def test_storage_basic_search_pin():
    # tests/test_storage.py:277-291: top-1 chunk_001, similarity > 0.8
    x = so.normalize_rows(X_STORAGE)
    res = so.storage_search(x, ROWS_STORAGE, np.array([.1, .2, .3, .4]))
    assert res[0][0] == 0 and res[0][1] > 0.8
    # restated scores quoted in SURVEY.md section 8c
    np.testing.assert_allclose([s for _, s in res], [0.9999999, 0.968864, 0.543220], atol=2e-6)


def test_storage_config_pin():
    # tests/test_storage.py:293-308: top_k=2, threshold 0.5
    x = so.normalize_rows(X_STORAGE)
    res = so.storage_search(x, ROWS_STORAGE, np.array([.1, .2, .3, .4]), top_k=2, similarity_threshold=0.5)
    assert len(res) <= 2 and all(s >= 0.5 for _, s in res)


@pytest.mark.parametrize("filters,count", [
    ({"project_name": "test_project"}, 2),          # :310-321
    ({"word_count": {"gte": 5}}, 2),                # :323-333
    ({"chunk_type": ["qa_pair", "code_block"]}, 2),  # :335-345
])
def test_storage_filter_count_pins(filters, count):
    x = so.normalize_rows(X_STORAGE)
    res = so.storage_search(x, ROWS_STORAGE, np.array([.1, .2, .3, .4]), filters=filters)
    assert len(res) == count


def test_matches_filters_truth_table():
    # tests/test_storage.py:617-647
    row = {"project_name": "test_project", "word_count": 10, "has_code": True, "chunk_type": "qa_pair"}
    assert so.matches_filters(row, {"project_name": "test_project"})
    assert not so.matches_filters(row, {"project_name": "other_project"})
    assert so.matches_filters(row, {"word_count": {"gte": 5}})
    assert so.matches_filters(row, {"word_count": {"lte": 15}})
    assert not so.matches_filters(row, {"word_count": {"gt": 10}})
    assert so.matches_filters(row, {"chunk_type": ["qa_pair", "code_block"]})
    assert not so.matches_filters(row, {"chunk_type": ["tool_usage"]})
    # unknown keys are ignored (src/storage.py:513-514; MCP related_to / same_session)
    assert so.matches_filters(row, {"related_to": "x", "same_session": True})


def test_empty_index_pin():
    # tests/test_storage.py:693-700
    assert so.storage_search(np.zeros((0, 4), np.float32), [], np.array([.1, .2, .3, .4])) == []


def test_integration_ranking_pin():
    # tests/test_integration.py:312-353: highly > somewhat > less, top sim > 0.9
    x = so.normalize_rows(np.array([[1, .9, .8, .7], [.8, .7, .6, .5], [.2, .3, .4, .5]], np.float32))
    rows = [dict(id=f"c{i}") for i in range(3)]
    res = so.storage_search(x, rows, np.array([1, .9, .8, .7]))
    assert [i for i, _ in res] == [0, 1, 2]
    assert res[0][1] > 0.9
    np.testing.assert_allclose([s for _, s in res], [1.0, 0.999218, 0.904762], atol=2e-6)


def test_environment_setup_pin():
    # tests/test_environment_setup.py:199-220: 10 random 128-d, search self k=5 -> I[0][0]==0
    rng = np.random.default_rng(0)
    v = rng.random((10, 128), dtype=np.float32)
    v = so.normalize_rows(v)
    D, I = so.flat_search(v, v[:1], 5)
    assert D.shape == (1, 5) and I.shape == (1, 5) and I[0][0] == 0


def test_project_filter_pins():
    # tests/test_project_filter.py:34-132 (substring, case-insensitive, combined)
    names = ["my-awesome-project", "my-awesome-project", "another-project", "MyAwesomeProject",
             "test-project"]
    has_code = [True, False, True, False, False]
    rows = [dict(id=f"c{i}", project_name=n, has_code=int(h)) for i, (n, h) in enumerate(zip(names, has_code))]
    rng = np.random.default_rng(1)
    x = so.normalize_rows(rng.random((5, 768), dtype=np.float32))
    q = rng.random(768, dtype=np.float32)
    count = lambda f: len(so.storage_search(x, rows, q, filters=f))
    assert count({"project_name": "awesome"}) == 3      # substring, case-insensitive
    assert count({"project_name": "AWESOME"}) == 3
    assert count({"project_name": "awesome", "has_code": True}) == 1
    assert count({"project_name": "project"}) == 5
    assert count({"project_name": "nonexistent"}) == 0


def test_c_restatement_agrees_with_numpy():
    rng = np.random.default_rng(2)
    x = so.normalize_rows(rng.standard_normal((5000, 768), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((7, 768), dtype=np.float32))
    for metric in (so.METRIC_IP, so.METRIC_L2):
        D, I = so.flat_search(x, q, 10, metric)
        Dc, Ic = so.flat_search_c(x, q, 10, metric)
        ok, why = so.compare_topk(D, I, Dc, Ic, tol=1e-5)
        assert ok, why
    mask = rng.random(5000) < 0.05
    D, I = so.flat_search(x, q, 10, mask=mask)
    Dc, Ic = so.flat_search_c(x, q, 10, mask_words=so.pack_mask(mask))
    ok, why = so.compare_topk(D, I, Dc, Ic, tol=1e-5)
    assert ok, why
    assert mask[I[I >= 0]].all()


def test_reference_result_is_prefix_of_prefilter():
    # SURVEY.md section 8a parity definition: R is a prefix of P
    rng = np.random.default_rng(3)
    n = 3000
    x = so.normalize_rows(rng.standard_normal((n, 64), dtype=np.float32))
    rows = [dict(id=f"c{i}", has_code=int(rng.random() < 0.05)) for i in range(n)]
    for _ in range(5):
        q = rng.standard_normal(64).astype(np.float32)
        R = so.storage_search(x, rows, q, top_k=10, filters={"has_code": True})
        P = so.prefilter_search(x, rows, q, top_k=10, filters={"has_code": True})
        assert [i for i, _ in P[:len(R)]] == [i for i, _ in R]
        assert len(P) == 10 and len(R) <= 10


def test_golden_fixture_roundtrip():
    g = np.load(GOLDEN / "search_small.npz")
    D, I = so.flat_search(g["x"], g["q"], int(g["k"]))
    np.testing.assert_array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], atol=1e-6)
    Dm, Im = so.flat_search(g["x"], g["q"], int(g["k"]), mask=g["mask"])
    np.testing.assert_array_equal(Im, g["I_masked"])
    Dl, Il = so.flat_search(g["x"], g["q"], int(g["k"]), metric=so.METRIC_L2)
    np.testing.assert_array_equal(Il, g["I_l2"])


def test_golden_768_regenerates():
    from oracle.make_golden import build_inputs
    g = np.load(GOLDEN / "search_768.npz")
    x, q, mask = build_inputs(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["nq"]))
    assert abs(float(x.astype(np.float64).sum()) - float(g["x_checksum"])) < 1e-6
    D, I = so.flat_search(x, q, int(g["k"]))
    np.testing.assert_array_equal(I, g["I"])
    # duplicates planted at rows 7, 100, n-148: equal scores must come out in id order
    assert list(I[0][:3]) == [7, 100, int(g["n"]) - 148]
