"""The reference's own unit tests for pieces of the host layer on the path, restated case by case (same inputs, same
expected values) against the Python mirror:

* data/src/test/scala/org/apache/predictionio/data/storage/BiMapSpec.scala:27-198 -> pio_b200.storage.BiMap
  (Option -> None, IllegalArgumentException -> ValueError, RDD[String] -> any iterable).

These are the only known-answer tests the reference holds for code on the ALS path (there are
none for the MLlib arithmetic itself, which is why the numeric oracle stays "parity unpinned")."""
import pytest

from pio_b200.storage import BiMap

KEYS = [1, 4, 6]
ORG_VALUES = [2, 5, 7]
ORG = dict(zip(KEYS, ORG_VALUES))


@pytest.fixture()
def bi():
    return BiMap(ORG)


def test_return_correct_values_for_each_key_of_original_map(bi):      # BiMapSpec.scala:40-44
    assert [bi(k) for k in KEYS] == ORG_VALUES


def test_get_returns_option(bi):                                       # :46-52
    assert [bi.get(k) for k in KEYS + [12345]] == ORG_VALUES + [None]


def test_get_or_else_returns_value_for_each_key(bi):                   # :54-58
    assert [bi.getOrElse(k, -1) for k in KEYS] == ORG_VALUES


def test_get_or_else_returns_default_for_invalid_key(bi):              # :60-66
    keys, defaults = [999, -1, -2], [1234, 5678, 987]
    assert [bi.getOrElse(k, d) for k, d in zip(keys, defaults)] == defaults


def test_contains(bi):                                                 # :68-74
    assert [bi.contains(k) for k in KEYS + [12345]] == [True, True, True, False]


def test_same_size_as_original_map(bi):                                # :76-78
    assert bi.size == len(ORG)


def test_take_2_returns_bimap_of_size_2(bi):                           # :80-82
    assert bi.take(2).size == 2


def test_to_map_and_to_seq_contain_same_elements(bi):                  # :84-90
    assert bi.toMap() == ORG
    assert sorted(bi.toSeq()) == sorted(ORG.items())


def test_inverse(bi):                                                  # :92-105
    assert [bi.inverse(v) for v in ORG_VALUES] == KEYS
    assert bi.inverse.size == len(ORG)
    assert bi.inverse.inverse is bi       # reference equality


def test_duplicated_values_are_rejected_at_construction():             # :108-113
    with pytest.raises(ValueError):
        BiMap({1: 2, 4: 7, 6: 7})


@pytest.mark.parametrize("keys", [
    {"a", "b", "foo", "bar"},                       # :117-130  Set[String]
    ["a", "b", "foo", "bar"],                       # :132-145  Array of unique strings / :163-177 RDD[String]
])
def test_string_long_and_string_int_from_unique_strings(keys):
    for make in (BiMap.stringLong, BiMap.stringInt):
        b = make(keys)
        assert sorted(b(k) for k in keys) == [0, 1, 2, 3]


def test_string_int_with_duplicated_strings():                         # :147-161, :179-195
    keys = ["a", "b", "foo", "bar", "a", "b", "x"]
    for make in (BiMap.stringLong, BiMap.stringInt):
        b = make(keys)
        distinct = list(dict.fromkeys(keys))
        # the RDD variant pins the index set to 0..n-1 over the distinct keys; the Array variant only promises distinct
        # indices - the mirror satisfies the stronger one
        assert sorted(b(k) for k in distinct) == [0, 1, 2, 3, 4]
        assert b.size == 5
