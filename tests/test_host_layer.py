"""CPU tests of the host-side mirror of the reference's DASE surface: BiMap (data/src/test/.../BiMapSpec.scala),
DataMap, the $set/$unset/$delete fold, engine.json -> Params extraction, Doer construction rules and the three
model-persistence modes pinned by core/src/test/.../controller/EngineTest.scala:66-185 (fake components carrying
ids, like SampleEngine.scala)."""
import datetime as dt
import json
from dataclasses import dataclass
from typing import Optional

import pytest

from pio_b200 import controller as c
from pio_b200 import storage as s
from pio_b200 import workflow as w


def test_bimap_semantics():
    b = s.BiMap.stringInt(["a", "b", "a", "c"])
    assert b.size == 3 and b("a") == 0 and b("c") == 2 and b.get("z") is None and b.getOrElse("z", -1) == -1
    assert b.inverse(1) == "b" and b.inverse.inverse is b and b.contains("b") and not b.contains("q")
    assert b.take(2).size == 2
    with pytest.raises(ValueError):
        s.BiMap({"a": 1, "b": 1})                     # duplicated values cannot be reversed (fails at construction)
    with pytest.raises(KeyError):
        b("nope")


def test_datamap_accessors():
    d = s.DataMap({"a": 1, "b": "x", "n": None})
    assert d.get("a", float) == 1.0 and d.getOpt("zz") is None and d.getOrElse("zz", 7) == 7
    with pytest.raises(s.DataMapException):
        d.get("zz")
    with pytest.raises(s.DataMapException):
        d.get("n")
    assert (d + s.DataMap({"a": 2})).get("a") == 2 and (d - ["a"]).keySet() == {"b", "n"}


def test_event_store_find_and_aggregate(tmp_path, monkeypatch):
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path))
    t0 = dt.datetime(2020, 1, 1, tzinfo=dt.timezone.utc)
    ev = lambda **k: dict(eventTime=(t0 + dt.timedelta(minutes=k.pop("m"))).isoformat(), **k)  # noqa: E731
    s.import_events("App", [
        ev(m=0, event="$set", entityType="item", entityId="i1", properties={"categories": ["c1"], "x": 1}),
        ev(m=1, event="$set", entityType="item", entityId="i2", properties={"categories": ["c2"]}),
        ev(m=2, event="$unset", entityType="item", entityId="i1", properties={"x": None}),
        ev(m=3, event="$delete", entityType="item", entityId="i2"),
        ev(m=4, event="rate", entityType="user", entityId="u1", targetEntityType="item", targetEntityId="i1",
           properties={"rating": 4.5}),
        ev(m=5, event="buy", entityType="user", entityId="u1", targetEntityType="item", targetEntityId="i2"),
        ev(m=6, event="view", entityType="user", entityId="u2"),
    ])
    assert len(s.PEventStore.find("App", entityType="user", eventNames=["rate", "buy"], targetEntityType="item")) == 2
    assert len(s.PEventStore.find("App", entityType="user", targetEntityType=None)) == 1   # Some(None): must be absent
    agg = dict(s.PEventStore.aggregateProperties("App", "item"))
    assert set(agg) == {"i1"} and agg["i1"].fields == {"categories": ["c1"]}
    assert s.PEventStore.aggregateProperties("App", "item", required=["x"]) == []
    with pytest.raises(FileNotFoundError):
        s.PEventStore.find("NoSuchApp")
    assert s.LEventStore.findByEntity("App", "user", "u1", limit=1)[0].event == "buy"      # latest first


# ---- fake DASE components with ids (SampleEngine.scala style) ------------------------------------
@dataclass
class DSP(c.Params):
    id: int
    en: Optional[int] = None


@dataclass
class AP(c.Params):
    id: int


class DS(c.PDataSource):
    def __init__(self, p: DSP):
        self.p = p

    def readTraining(self, sc):
        return ("td", self.p.id)


class Prep(c.PPreparator):
    def prepare(self, sc, td):
        return ("pd", td)


class Model:
    def __init__(self, id, pd):
        self.id, self.pd = id, pd

    def __eq__(self, o):
        return type(o) is type(self) and (o.id, o.pd) == (self.id, self.pd)


class PModel(Model, c.PersistentModel):
    store = {}

    def save(self, id, params, sc):
        PModel.store[id] = self
        return True

    @classmethod
    def apply(cls, id, params, sc):
        return PModel.store[id]


class PAlgo0(c.PAlgorithm):          # parallel model that is not persistent -> Unit -> re-train at deploy
    def __init__(self, p: AP):
        self.p = p

    def train(self, sc, pd):
        return Model(self.p.id, pd)

    def predict(self, model, query):
        return (self.p.id, model.id, query)


class PAlgo1(PAlgo0):                # PersistentModel -> manifest
    def train(self, sc, pd):
        return PModel(self.p.id, pd)


class LAlgo0(c.P2LAlgorithm):        # local model -> stored as is
    def __init__(self, p: AP):
        self.p = p

    def train(self, sc, pd):
        return Model(self.p.id, pd)

    def predict(self, model, query):
        return (self.p.id, model.id, query)


class Serv(c.LServing):
    def serve(self, query, predictions):
        return predictions


def _engine():
    return c.Engine(DS, Prep, {"PAlgo0": PAlgo0, "PAlgo1": PAlgo1, "LAlgo0": LAlgo0}, Serv)


def test_engine_json_to_params_and_doer():
    e = _engine()
    ep = e.jValueToEngineParams({"datasource": {"params": {"id": 3, "ignored": 1}},
                                 "algorithms": [{"name": "PAlgo1", "params": {"id": 5}},
                                                {"name": "LAlgo0", "params": {"id": 6}}]})
    assert ep.dataSourceParams == ("", DSP(3, None)) and [n for n, _ in ep.algorithmParamsList] == ["PAlgo1", "LAlgo0"]
    with pytest.raises(ValueError):
        e.jValueToEngineParams({"datasource": {"params": {}}, "algorithms": []})          # missing required id
    with pytest.raises(ValueError):
        e.jValueToEngineParams({"datasource": {"params": {"id": 1}}, "algorithms": [{"name": "nope", "params": {}}]})
    assert isinstance(c.Doer.apply(Prep, c.EmptyParams()), Prep)                            # zero-arg fallback
    assert c.Doer.apply(DS, DSP(9)).p.id == 9


def test_train_persistence_modes_and_prepare_deploy():
    """EngineTest.scala:66-185 -- Unit / PersistentModelManifest / model itself; then prepareDeploy restores all."""
    e = _engine()
    ep = c.EngineParams(dataSourceParams=("", DSP(1)), preparatorParams=("", c.EmptyParams()),
                        algorithmParamsList=[("PAlgo0", AP(2)), ("PAlgo1", AP(3)), ("LAlgo0", AP(4))],
                        servingParams=("", c.EmptyParams()))
    persisted = e.train(None, ep, "inst", w.WorkflowParams())
    pd = ("pd", ("td", 1))
    assert persisted[0] is c.Unit
    assert isinstance(persisted[1], c.PersistentModelManifest) and persisted[1].className.endswith("PModel")
    assert persisted[2] == Model(4, pd)
    import pickle
    restored = e.prepareDeploy(None, ep, "inst", pickle.loads(pickle.dumps(persisted)))
    assert restored[0] == Model(2, pd) and restored[1] == PModel(3, pd) and restored[2] == Model(4, pd)


def test_stop_after_read_and_prepare():
    e = _engine()
    ep = c.EngineParams(dataSourceParams=("", DSP(1)), algorithmParamsList=[("LAlgo0", AP(4))])
    with pytest.raises(c.StopAfterReadInterruption):
        e.train(None, ep, "i", w.WorkflowParams(stopAfterRead=True))
    with pytest.raises(c.StopAfterPrepareInterruption):
        e.train(None, ep, "i", w.WorkflowParams(stopAfterPrepare=True))


def test_create_workflow_registry_and_query_server(tmp_path, monkeypatch):
    monkeypatch.setenv("PIO_MODELDATA_DIR", str(tmp_path))
    variant = tmp_path / "engine.json"
    variant.write_text(json.dumps({"id": "default", "engineFactory": "tests.test_host_layer:_engine",
                                   "datasource": {"params": {"id": 1}},
                                   "algorithms": [{"name": "LAlgo0", "params": {"id": 7}}],
                                   "sparkConf": {"spark": {"executor": {"extraJavaOptions": "x"}}}}))
    inst = w.CreateWorkflow.main(["--engine-id", "E", "--engine-version", "1", "--engine-variant", str(variant),
                                  "--some-unknown-flag", "tolerated"])
    assert inst.status == "COMPLETED" and inst.sparkConf == {"spark.executor.extraJavaOptions": "x"}
    qs = w.deploy(engineId="E", engineVersion="1", engineVariant="default")
    assert qs.query({"q": 1}) == [[7, 7, {"q": 1}]]
    with pytest.raises(RuntimeError):
        w.deploy(engineId="other", engineVersion="1")
