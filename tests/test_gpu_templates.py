"""GPU tests (-m gpu) of the drop-in path a template user takes: engine.json -> CreateWorkflow -> Engine.train ->
ALSAlgorithm.train -> native ALS -> persisted model -> deploy -> query, compared with the oracle by STRING id
(BiMap index order is not part of the contract, SURVEY hard part 5)."""
import datetime as dt
import json

import numpy as np
import pytest

from pio_b200 import storage as s
from pio_b200 import synth
from pio_b200 import workflow as w

pytestmark = pytest.mark.gpu


def _events(nu, ni, nnz, seed, implicit=False):
    u, i, r = synth.synth_ratings(nu, ni, nnz, seed=seed, implicit=implicit)
    t0 = dt.datetime(2021, 1, 1, tzinfo=dt.timezone.utc)
    evs = []
    for e in range(nnz):
        if implicit:
            for _ in range(int(r[e])):
                evs.append(dict(event="view", entityType="user", entityId=f"u{u[e]}", targetEntityType="item",
                                targetEntityId=f"i{i[e]}", eventTime=(t0 + dt.timedelta(seconds=len(evs))).isoformat()))
        elif e % 3 == 0:
            evs.append(dict(event="buy", entityType="user", entityId=f"u{u[e]}", targetEntityType="item",
                            targetEntityId=f"i{i[e]}", eventTime=(t0 + dt.timedelta(seconds=e)).isoformat()))
        else:
            evs.append(dict(event="rate", entityType="user", entityId=f"u{u[e]}", targetEntityType="item",
                            targetEntityId=f"i{i[e]}", properties={"rating": float(r[e])},
                            eventTime=(t0 + dt.timedelta(seconds=e)).isoformat()))
    return evs


def test_recommendation_template_end_to_end(tmp_path, monkeypatch, oracle):
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path / "events"))
    monkeypatch.setenv("PIO_MODELDATA_DIR", str(tmp_path / "models"))
    evs = _events(400, 80, 5000, seed=3)
    s.import_events("MyApp1", evs)
    variant = tmp_path / "engine.json"
    variant.write_text(json.dumps({
        "id": "default", "description": "Default settings",
        "engineFactory": "pio_b200.templates.recommendation.RecommendationEngine",
        "datasource": {"params": {"appName": "MyApp1"}},
        "algorithms": [{"name": "als", "params": {"rank": 10, "numIterations": 20, "lambda": 0.01, "seed": 3}}]}))
    inst = w.CreateWorkflow.main(["--engine-id", "rec", "--engine-version", "1", "--engine-variant", f"file:{variant}"])
    assert inst.status == "COMPLETED"
    server = w.deploy(inst.id)          # reloads the PersistentModel (ALSModel.apply) onto the GPU
    model = server.models[0]

    # oracle on the same ratings, same id maps, same initial factors
    users = [e["entityId"] for e in evs]
    items = [e["targetEntityId"] for e in evs]
    um, im = s.BiMap.stringInt(users), s.BiMap.stringInt(items)
    assert um.toMap() == model.userStringIntMap.toMap() and im.toMap() == model.itemStringIntMap.toMap()
    u = np.array([um(x) for x in users], np.int32)
    i = np.array([im(x) for x in items], np.int32)
    r = np.array([4.0 if e["event"] == "buy" else e["properties"]["rating"] for e in evs], np.float32)
    u0, i0 = synth.synth_init_factors(um.size, 10, 3, 0), synth.synth_init_factors(im.size, 10, 3, 1)
    ouf, oitf, ouh, oih = oracle.als_train(um.size, im.size, u, i, r, 10, 20, 0.01, False, 1.0, u0, i0)
    assert np.linalg.norm(model.userFeatures - ouf) / np.linalg.norm(ouf) <= 1e-4
    assert np.linalg.norm(model.productFeatures - oitf) / np.linalg.norm(oitf) <= 1e-4

    # the reference's only ALS assertion: 4 itemScores for num=4 (quickstart_test.py:163-167)
    res = server.query({"user": "u1", "num": 4})
    assert len(res["itemScores"]) == 4
    oi, os_, _ = oracle.recommend(model.userFeatures, model.userHas, model.productFeatures, model.productHas,
                                  np.array([um("u1")], np.int32), 4)
    assert [x["item"] for x in res["itemScores"]] == [im.inverse(int(t)) for t in oi[0]]
    bl = res["itemScores"][0]["item"]
    res2 = server.query({"user": "u1", "num": 4, "blackList": [bl]})
    assert bl not in [x["item"] for x in res2["itemScores"]] and len(res2["itemScores"]) == 4
    assert server.query({"user": "nobody", "num": 4}) == {"itemScores": []}


def test_similarproduct_template(tmp_path, monkeypatch, oracle):
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path / "events"))
    monkeypatch.setenv("PIO_MODELDATA_DIR", str(tmp_path / "models"))
    nu, ni = 200, 50
    t0 = dt.datetime(2021, 1, 1, tzinfo=dt.timezone.utc).isoformat()
    sets = [dict(event="$set", entityType="user", entityId=f"u{k}", eventTime=t0) for k in range(nu)]
    sets += [dict(event="$set", entityType="item", entityId=f"i{k}", eventTime=t0,
                  properties={"categories": ["even" if k % 2 == 0 else "odd"]}) for k in range(ni + 5)]  # 5 never viewed
    s.import_events("Sim", sets + _events(nu, ni, 1500, seed=5, implicit=True))
    from pio_b200.templates import similarproduct as sp
    eng = sp.SimilarProductEngine().apply()
    ep = eng.jValueToEngineParams({"datasource": {"params": {"appName": "Sim"}},
                                   "algorithms": [{"name": "als", "params": {"rank": 8, "numIterations": 5,
                                                                             "lambda": 0.01, "seed": 3}}]})
    sc = w.WorkflowContext()
    models = eng.prepareDeploy(sc, ep, "simtest", eng.train(sc, ep, "simtest"))
    m = models[0]
    algo = sp.ALSAlgorithm(ep.algorithmParamsList[0][1])
    res = algo.predict(m, sp.Query(items=["i1", "i3"], num=5))
    assert len(res.itemScores) == 5 and not ({"i1", "i3"} & {x.item for x in res.itemScores})
    q = np.array(sorted([m.itemStringIntMap("i1"), m.itemStringIntMap("i3")]), np.int32)
    oi, os_, oc = oracle.similar(m.mf.productFeatures, m.mf.productHas, q, 5)
    assert [x.item for x in res.itemScores] == [m.itemIntStringMap(int(t)) for t in oi[:oc]]
    res = algo.predict(m, sp.Query(items=["i1"], num=50, categories={"even"}, blackList={"i2"}))
    assert res.itemScores and all(int(x.item[1:]) % 2 == 0 and x.item != "i2" for x in res.itemScores)
    assert all(x.score > 0 for x in res.itemScores)
    assert algo.predict(m, sp.Query(items=["i52"], num=3)).itemScores == []     # item without a factor


def test_classification_template(tmp_path, monkeypatch, oracle):
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path / "events"))
    rng = np.random.default_rng(0)
    t0 = dt.datetime(2021, 1, 1, tzinfo=dt.timezone.utc).isoformat()
    rows = [(int(rng.integers(0, 4)), *[int(v) for v in rng.integers(0, 10, 3)]) for _ in range(500)]
    s.import_events("Cls", [dict(event="$set", entityType="user", entityId=f"u{k}", eventTime=t0,
                                 properties={"plan": p, "attr0": a, "attr1": b, "attr2": c2})
                            for k, (p, a, b, c2) in enumerate(rows)])
    from pio_b200.templates import classification as cl
    eng = cl.ClassificationEngine().apply()
    ep = eng.jValueToEngineParams({"datasource": {"params": {"appName": "Cls"}},
                                   "algorithms": [{"name": "naive", "params": {"lambda": 1.0}}]})
    sc = w.WorkflowContext()
    model = eng.train(sc, ep, "cls")[0]
    y = np.array([r[0] for r in rows], np.int32)
    x = np.array([r[1:] for r in rows], np.float32)
    opi, oth = oracle.nb_train(y, x, 4, 1.0)
    assert np.array_equal(model.pi, opi) and np.array_equal(model.theta, oth)
    algo = cl.NaiveBayesAlgorithm(ep.algorithmParamsList[0][1])
    pred = [algo.predict(model, cl.Query(*map(float, r[1:]))).label for r in rows[:50]]
    assert pred == [float(v) for v in oracle.nb_predict(x[:50], opi, oth)]


def test_ecommerce_template(tmp_path, monkeypatch, oracle):
    """ECommAlgorithm: explicit ALS.train on rate events where the LATEST rating of a (user,item) pair wins
    (genMLlibRating, ECommAlgorithm.scala:163-203) -- the reduceByKey runs on the GPU (dedup = keep_last)."""
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path / "events"))
    monkeypatch.setenv("PIO_MODELDATA_DIR", str(tmp_path / "models"))
    nu, ni = 120, 40
    t0 = dt.datetime(2021, 1, 1, tzinfo=dt.timezone.utc)
    sets = [dict(event="$set", entityType="user", entityId=f"u{k}", eventTime=t0.isoformat()) for k in range(nu)]
    sets += [dict(event="$set", entityType="item", entityId=f"i{k}", eventTime=t0.isoformat(),
                  properties={"categories": ["c%d" % (k % 3)]}) for k in range(ni)]
    rng = np.random.default_rng(5)
    evs = []
    for e in range(2500):     # many repeated pairs with different times / ratings
        evs.append(dict(event="rate", entityType="user", entityId=f"u{rng.integers(nu)}", targetEntityType="item",
                        targetEntityId=f"i{rng.integers(ni)}", properties={"rating": float(rng.integers(1, 6))},
                        eventTime=(t0 + dt.timedelta(seconds=int(rng.integers(0, 100000)))).isoformat()))
    for e in range(300):
        evs.append(dict(event="buy", entityType="user", entityId=f"u{rng.integers(nu)}", targetEntityType="item",
                        targetEntityId=f"i{rng.integers(ni)}", eventTime=(t0 + dt.timedelta(seconds=e)).isoformat()))
    s.import_events("Shop", sets + evs)
    from pio_b200.templates import ecommerce as ec
    eng = ec.ECommerceRecommendationEngine().apply()
    ep = eng.jValueToEngineParams({"datasource": {"params": {"appName": "Shop"}},
                                   "algorithms": [{"name": "ecomm", "params": {
                                       "appName": "Shop", "unseenOnly": True, "seenEvents": ["buy"],
                                       "similarEvents": ["view"], "rank": 8, "numIterations": 6, "lambda": 0.05,
                                       "seed": 3}}]})
    sc = w.WorkflowContext()
    m = eng.prepareDeploy(sc, ep, "shop", eng.train(sc, ep, "shop"))[0]
    # oracle with host-side latest-wins preparation
    rates = [e for e in evs if e["event"] == "rate"]
    um, im = m.userStringIntMap, m.itemStringIntMap
    u = np.array([um(e["entityId"]) for e in rates], np.int32)
    i = np.array([im(e["targetEntityId"]) for e in rates], np.int32)
    r = np.array([e["properties"]["rating"] for e in rates], np.float32)
    ts = np.array([int(dt.datetime.fromisoformat(e["eventTime"]).timestamp() * 1000) for e in rates], np.int64)
    uu, ii, rr = oracle.dedup_coo(u, i, r, "keep_last", ts)
    u0, i0 = synth.synth_init_factors(um.size, 8, 3, 0), synth.synth_init_factors(im.size, 8, 3, 1)
    ouf, oitf, ouh, oih = oracle.als_train(um.size, im.size, uu, ii, rr, 8, 6, 0.05, False, 1.0, u0, i0)
    assert np.linalg.norm(m.mf.userFeatures - ouf) / np.linalg.norm(ouf) <= 1e-4
    assert np.linalg.norm(m.mf.productFeatures - oitf) / np.linalg.norm(oitf) <= 1e-4
    algo = ec.ECommAlgorithm(ep.algorithmParamsList[0][1])
    known = next(e["entityId"] for e in rates)
    bought = {e["targetEntityId"] for e in evs if e["event"] == "buy" and e["entityId"] == known}
    res = algo.predict(m, ec.Query(user=known, num=5))
    assert res.itemScores and all(x.score > 0 for x in res.itemScores)
    assert not (bought & {x.item for x in res.itemScores})                    # unseenOnly
    res = algo.predict(m, ec.Query(user=known, num=40, categories={"c1"}))
    assert all(int(x.item[1:]) % 3 == 1 for x in res.itemScores)
    res = algo.predict(m, ec.Query(user="stranger", num=3))                   # unknown user, no recent views -> popularity
    counts = {}
    for e in evs:
        if e["event"] == "buy":
            counts[e["targetEntityId"]] = counts.get(e["targetEntityId"], 0) + 1
    assert [x.score for x in res.itemScores] == sorted(counts.values(), reverse=True)[:3]
