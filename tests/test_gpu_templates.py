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

    # CooccurrenceAlgorithm of the same engine (CooccurrenceAlgorithm.scala:44-175): counting on the GPU, predict on the host
    ep2 = eng.jValueToEngineParams({"datasource": {"params": {"appName": "Sim"}},
                                    "algorithms": [{"name": "cooccurrence", "params": {"n": 10}}]})
    cm = eng.train(sc, ep2, "cooc")[0]
    calgo = sp.CooccurrenceAlgorithm(ep2.algorithmParamsList[0][1])
    td = sp.DataSource(ep2.dataSourceParams[1]).readTraining(sc)
    per_user = {}
    for e in td.viewEvents:
        per_user.setdefault(e.user, set()).add(cm.itemStringIntMap(e.item))
    want = {}
    for its in per_user.values():
        for a in its:
            for b in its:
                if a != b:
                    want.setdefault(a, {}).setdefault(b, 0)
                    want[a][b] += 1
    for it in (cm.itemStringIntMap("i1"), cm.itemStringIntMap("i3"), cm.itemStringIntMap("i20")):
        exp = sorted(want.get(it, {}).items(), key=lambda kv: (-kv[1], kv[0]))[:10]
        assert cm.topCooccurrences(it) == exp
    res = calgo.predict(cm, sp.Query(items=["i1", "i3"], num=5, blackList={"i0"}))
    assert len(res.itemScores) == 5 and not ({"i0", "i1", "i3"} & {x.item for x in res.itemScores})
    assert [x.score for x in res.itemScores] == sorted((x.score for x in res.itemScores), reverse=True)
    assert calgo.predict(cm, sp.Query(items=["i52"], num=3)).itemScores == []    # never viewed


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
    # adjust-score variant: a $set of the constraint entity "weightedItems" multiplies the scores of the listed items
    # (adjust-score ECommAlgorithm.scala:258-266,490-497); weight 0 removes an item (score > 0 filter), a large weight
    # promotes it
    top = [x.item for x in res.itemScores]
    s.import_events("Shop", [dict(event="$set", entityType="constraint", entityId="weightedItems",
                                  eventTime=(t0 + dt.timedelta(days=30)).isoformat(),
                                  properties={"weights": [{"items": [top[0]], "weight": 0.0},
                                                          {"items": [top[-1]], "weight": 1000.0}]})])
    res_w = algo.predict(m, ec.Query(user=known, num=5))
    assert top[0] not in [x.item for x in res_w.itemScores] and res_w.itemScores[0].item == top[-1]
    uidx = m.userStringIntMap(known)
    wv = np.ones(m.itemStringIntMap.size)
    wv[m.itemStringIntMap(top[0])] = 0.0
    wv[m.itemStringIntMap(top[-1])] = 1000.0
    mask = algo._mask(m, ec.Query(user=known, num=5), {b for b in (m.itemStringIntMap.get(x) for x in algo.genBlackList(
        ec.Query(user=known, num=5))) if b is not None})
    oi, os_, oc = oracle.recommend(m.mf.userFeatures, m.mf.userHas, m.mf.productFeatures, m.mf.productHas,
                                   np.array([uidx], np.int32), 5, mask, wv)
    assert [x.item for x in res_w.itemScores] == [m.itemIntStringMap(int(t)) for t, sc_ in zip(oi[0], os_[0]) if sc_ > 0]
    # predictSimilar (unknown user with recent views): the recently viewed items stay candidates
    # (train-with-rate-event ECommAlgorithm.scala:492-525 has no "not a query item" rule)
    s.import_events("Shop", [dict(event="view", entityType="user", entityId="newcomer", targetEntityType="item",
                                  targetEntityId=top[1], eventTime=(t0 + dt.timedelta(days=31)).isoformat())])
    res_s = algo.predict(m, ec.Query(user="newcomer", num=3))
    assert top[1] in [x.item for x in res_s.itemScores]      # with similarproduct's rule it could never appear
    res = algo.predict(m, ec.Query(user=known, num=40, categories={"c1"}))
    assert all(int(x.item[1:]) % 3 == 1 for x in res.itemScores)
    res = algo.predict(m, ec.Query(user="stranger", num=3))                   # unknown user, no recent views -> popularity
    counts = {}
    for e in evs:
        if e["event"] == "buy":
            counts[e["targetEntityId"]] = counts.get(e["targetEntityId"], 0) + 1
    # predictDefault multiplies the popularity count by the item weight as well (adjust-score :505-533)
    wmap = {top[0]: 0.0, top[-1]: 1000.0}
    assert [x.score for x in res.itemScores] == sorted((c * wmap.get(it, 1.0) for it, c in counts.items()), reverse=True)[:3]


def test_recommendation_evaluation_k_fold_on_gpu(tmp_path, monkeypatch, oracle):
    """`pio eval` of the recommendation template (Evaluation.scala:64-110): every fold of every engine-params set is one
    ALS training + one batched top-N call on the device (3 x 3 sets x 5 folds = 45 trainings).  One fold is recomputed with
    the oracle (same split, same maps, same seed) and must give the same Precision@K."""
    from pio_b200 import evaluation as ev
    from pio_b200.templates import recommendation as rec
    monkeypatch.setenv("PIO_EVENTDATA_DIR", str(tmp_path / "events"))
    monkeypatch.setenv("PIO_MODELDATA_DIR", str(tmp_path / "models"))
    evs = _events(300, 60, 6000, seed=5)
    s.import_events("MyApp1", evs)
    gen = rec.EngineParamsList(appName="MyApp1")
    assert len(gen.engineParamsList) == 9
    res = ev.run_evaluation(rec.RecommendationEvaluation(), gen)
    assert len(res.engineParamsScores) == 9 and res.metricHeader == "Precision@K (k=10, threshold=4.0)"
    assert len(res.otherMetricHeaders) == 5
    scores = [sc.score for _, sc in res.engineParamsScores]
    assert all(0.0 <= x <= 1.0 for x in scores) and res.bestScore.score == max(scores)
    assert res.bestIdx == scores.index(max(scores))                       # the first maximum wins
    pos = [sc.otherScores[0] for _, sc in res.engineParamsScores]
    assert max(pos) == min(pos) > 0                                        # PositiveCount does not depend on the model

    # fold 0 of the first parameter set (rank 5, 1 iteration) through the oracle
    ep = gen.engineParamsList[0]
    ds = rec.DataSource(ep.dataSourceParams[1])
    td, _, qas = ds.readEval(None)[0]
    users = [r.user for r in td.ratings]
    items = [r.item for r in td.ratings]
    um, im = s.BiMap.stringInt(users), s.BiMap.stringInt(items)
    u = np.array([um(x) for x in users], np.int32)
    i = np.array([im(x) for x in items], np.int32)
    r = np.array([x.rating for x in td.ratings], np.float32)
    u0, i0 = synth.synth_init_factors(um.size, 5, 3, 0), synth.synth_init_factors(im.size, 5, 3, 1)
    ouf, oitf, ouh, oih = oracle.als_train(um.size, im.size, u, i, r, 5, 1, 0.01, False, 1.0, u0, i0)
    metric = rec.PrecisionAtK(10, 4.0)
    vals = []
    for q, a in qas:
        uidx = um.get(q.user)
        if uidx is None:
            p = rec.PredictedResult([])
        else:
            oi, os_, oc = oracle.recommend(ouf, ouh, oitf, oih, np.array([uidx], np.int32), q.num)
            p = rec.PredictedResult([rec.ItemScore(im.inverse(int(oi[0, t])), float(os_[0, t])) for t in range(int(oc[0]))])
        v = metric.calculate_one(q, p, a)
        if v is not None:
            vals.append(v)
    want = sum(vals) / len(vals)
    engine = rec.RecommendationEvaluation.engine
    got = metric.calculate(None, [engine.eval(w.WorkflowContext(mode="Evaluation"), ep)[0]])
    assert abs(got - want) <= 1e-9, (got, want)
