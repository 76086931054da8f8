"""The reference's own unit tests for pieces of the host layer on the path, restated case by case (same inputs, same
expected values) against the Python mirror:

* data/src/test/scala/org/apache/predictionio/data/storage/BiMapSpec.scala:27-198 -> pio_b200.storage.BiMap
  (Option -> None, IllegalArgumentException -> ValueError, RDD[String] -> any iterable).

* data/src/test/scala/org/apache/predictionio/data/storage/DataMapSpec.scala:24-207 -> pio_b200.storage.DataMap
  (get[T] / getOpt[T] / extract[T]; case classes -> dataclasses; the spec's JSON literals lack two commas that json4s
  tolerates - they are restored here).
* data/src/test/scala/org/apache/predictionio/data/storage/LEventAggregatorSpec.scala:28-104 with the events of
  TestEvents.scala:25-125 -> pio_b200.storage.LEventAggregator (the $set / $unset / $delete fold behind
  PEventStore.aggregateProperties, which the similarproduct / ecommerce DataSources call).

These are the only known-answer tests the reference holds for code on the ALS path (there are
none for the MLlib arithmetic itself, which is why the numeric oracle stays "parity unpinned")."""
import datetime as dt
from dataclasses import dataclass, replace
from typing import List, Optional, Set

import pytest

from pio_b200.storage import BiMap, DataMap, Event, LEventAggregator, PropertyMap

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


# ---------------------------------------------------------------------------------------------------------------
# DataMapSpec.scala
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Context:                      # DataMapSpec.scala:214-220
    ip: str
    prop1: Optional[float]
    prop2: Optional[str]
    prop3: Optional[int]
    prop4: List[int]


@dataclass
class BasicProperty:                # :222-229
    prop1: int
    prop2: str
    prop3: List[int]
    prop4: bool
    prop5: List[str]
    prop6: float


@dataclass
class OptionProperty:               # :231-238
    prop1: Optional[int]
    prop2: Optional[str]
    prop3: Optional[List[int]]
    prop4: Optional[bool]
    prop5: Optional[List[str]]
    prop6: Optional[float]


@dataclass
class MultiLevelProperty:           # :240-244
    context: Context
    anotherPropertyA: float
    anotherPropertyB: bool


BASIC_JSON = """{"prop1": 1, "prop2": "value2", "prop3": [1, 2, 3], "prop4": true, "prop5": ["a", "b", "c", "c"],
                 "prop6": 4.56}"""
MULTI_JSON = """{"context": {"ip": "1.23.4.56", "prop1": 2.345, "prop2": "value1", "prop4": [1, 2, 3]},
                 "anotherPropertyA": 4.567, "anotherPropertyB": false}"""
CONTEXT = Context(ip="1.23.4.56", prop1=2.345, prop2="value1", prop3=None, prop4=[1, 2, 3])


def test_datamap_typed_getters():                                      # DataMapSpec.scala:24-75
    p = DataMap.fromJson(BASIC_JSON)
    assert p.get("prop1", int) == 1 and p.getOpt("prop1", int) == 1
    assert p.get("prop2", str) == "value2" and p.getOpt("prop2", str) == "value2"
    assert p.get("prop3", List[int]) == [1, 2, 3] and p.getOpt("prop3", List[int]) == [1, 2, 3]
    assert p.get("prop4", bool) is True and p.getOpt("prop4", bool) is True
    assert p.get("prop5", List[str]) == ["a", "b", "c", "c"]
    assert p.get("prop5", Set[str]) == {"a", "b", "c"} and p.getOpt("prop5", Set[str]) == {"a", "b", "c"}
    assert p.get("prop6", float) == 4.56 and p.getOpt("prop6", float) == 4.56
    assert p.getOpt("prop9999", int) is None


def test_datamap_multi_level_data():                                   # :77-113
    p = DataMap.fromJson(MULTI_JSON)
    assert p.get("context", Context) == CONTEXT
    assert p.getOpt("context999", Context) is None
    assert p.get("anotherPropertyA", float) == 4.567
    assert p.get("anotherPropertyB", bool) is False


def test_datamap_extract():                                            # :115-207
    assert DataMap.fromJson(BASIC_JSON).extract(BasicProperty) == BasicProperty(
        prop1=1, prop2="value2", prop3=[1, 2, 3], prop4=True, prop5=["a", "b", "c", "c"], prop6=4.56)
    assert DataMap.fromJson("{}").extract(OptionProperty) == OptionProperty(None, None, None, None, None, None)
    some = DataMap.fromJson('{"prop1": 1, "prop5": ["a", "b", "c", "c"], "prop6": 4.56}').extract(OptionProperty)
    assert some == OptionProperty(prop1=1, prop2=None, prop3=None, prop4=None, prop5=["a", "b", "c", "c"], prop6=4.56)
    assert DataMap.fromJson(MULTI_JSON).extract(MultiLevelProperty) == MultiLevelProperty(
        context=CONTEXT, anotherPropertyA=4.567, anotherPropertyB=False)


# ---------------------------------------------------------------------------------------------------------------
# TestEvents.scala + LEventAggregatorSpec.scala
# ---------------------------------------------------------------------------------------------------------------
def _t(millis):
    return dt.datetime.fromtimestamp(millis / 1000.0, tz=dt.timezone.utc)


DAY = dt.timedelta(days=1)
U1_BASE, U2_BASE = _t(654321), _t(6543210)                             # TestEvents.scala:25-26
u1e1 = Event(event="$set", entityType="user", entityId="u1",
             properties=DataMap({"a": 1, "b": "value2", "d": [1, 2, 3]}), eventTime=U1_BASE)
u1e2 = replace(u1e1, properties=DataMap({"a": 2}), eventTime=U1_BASE + DAY)
u1e3 = replace(u1e1, properties=DataMap({"b": "value4"}), eventTime=U1_BASE + 2 * DAY)
u1e4 = replace(u1e1, event="$unset", properties=DataMap({"b": None}), eventTime=U1_BASE + 3 * DAY)
u1e5 = replace(u1e1, properties=DataMap({"e": "new"}), eventTime=U1_BASE + 4 * DAY)
U1_LAST = U1_BASE + 4 * DAY
U1 = {"a": 2, "d": [1, 2, 3], "e": "new"}                              # :66
u1ed = replace(u1e1, event="$delete", properties=DataMap(), eventTime=U1_BASE + 5 * DAY)
u2e1 = Event(event="$set", entityType="user", entityId="u2",
             properties=DataMap({"a": 21, "b": "value12", "d": [7, 5, 6]}), eventTime=U2_BASE)
u2e2 = replace(u2e1, event="$unset", properties=DataMap({"a": None}), eventTime=U2_BASE + DAY)
u2e3 = replace(u2e1, properties=DataMap({"b": "value9", "g": "new11"}), eventTime=U2_BASE + 2 * DAY)
U2_LAST = U2_BASE + 2 * DAY
U2 = {"b": "value9", "d": [7, 5, 6], "g": "new11"}                     # :101


def test_aggregate_two_entities():                                     # LEventAggregatorSpec.scala:30-54
    events = [u1e5, u2e2, u1e3, u1e1, u2e3, u2e1, u1e4, u1e2]
    result = LEventAggregator.aggregateProperties(iter(events))
    assert {k: v.fields for k, v in result.items()} == {"u1": U1, "u2": U2}
    assert result == {"u1": PropertyMap(U1, U1_BASE, U1_LAST), "u2": PropertyMap(U2, U2_BASE, U2_LAST)}


def test_aggregate_deleted_entity():                                   # :57-66
    events = [u1e5, u2e2, u1e3, u1ed, u1e1, u2e3, u2e1, u1e4, u1e2]
    assert LEventAggregator.aggregateProperties(iter(events)) == {"u2": PropertyMap(U2, U2_BASE, U2_LAST)}


def test_aggregate_single_entity():                                    # :70-103
    events = [u1e5, u1e3, u1e1, u1e4, u1e2]
    result = LEventAggregator.aggregatePropertiesSingle(iter(events))
    assert result.fields == U1 and result == PropertyMap(U1, U1_BASE, U1_LAST)
    # the delete event in the middle of the input; it is the latest in event time
    assert LEventAggregator.aggregatePropertiesSingle(iter([u1e4, u1e2, u1ed, u1e3, u1e1, u1e5])) is None
