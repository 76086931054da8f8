"""scala-parallel-similarproduct (multi-events-multi-algos: ALSAlgorithm on view events, LikeAlgorithm on
like/dislike events; Serving merges by standardised score).

Mirrors examples/scala-parallel-similarproduct/multi-events-multi-algos/src/main/scala/:
  Engine.scala, DataSource.scala, Preparator.scala, ALSAlgorithm.scala:33-263, LikeAlgorithm.scala:37-115,
  Serving.scala:29-69.  CooccurrenceAlgorithm is not ALS and is out of scope.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Set

import numpy as np

from ..controller import (Engine, EngineFactory, LServing, P2LAlgorithm, Params, PDataSource, PersistentModel,
                          PPreparator)
from ..mllib import ALS, MatrixFactorizationModel
from ..storage import BiMap, PEventStore


@dataclass
class Query:
    items: List[str]
    num: int
    categories: Optional[Set[str]] = None
    categoryBlackList: Optional[Set[str]] = None
    whiteList: Optional[Set[str]] = None
    blackList: Optional[Set[str]] = None


@dataclass
class ItemScore:
    item: str
    score: float


@dataclass
class PredictedResult:
    itemScores: List[ItemScore]


@dataclass
class DataSourceParams(Params):
    appName: str


@dataclass
class User:
    pass


@dataclass
class Item:
    categories: Optional[List[str]] = None


@dataclass
class ViewEvent:
    user: str
    item: str
    t: int


@dataclass
class LikeEvent:
    user: str
    item: str
    t: int
    like: bool


class TrainingData:
    def __init__(self, users: Dict[str, User], items: Dict[str, Item], viewEvents: List[ViewEvent],
                 likeEvents: List[LikeEvent]):
        self.users, self.items, self.viewEvents, self.likeEvents = users, items, viewEvents, likeEvents


PreparedData = TrainingData


class DataSource(PDataSource):
    def __init__(self, dsp: DataSourceParams):
        self.dsp = dsp

    def readTraining(self, sc) -> TrainingData:
        users = {k: User() for k, _ in PEventStore.aggregateProperties(self.dsp.appName, "user", sc=sc)}
        items = {k: Item(categories=pm.getOpt("categories"))
                 for k, pm in PEventStore.aggregateProperties(self.dsp.appName, "item", sc=sc)}
        evs = PEventStore.find(self.dsp.appName, entityType="user", eventNames=["view", "like", "dislike"],
                               targetEntityType="item", sc=sc)
        views, likes = [], []
        for e in evs:
            t = int(e.eventTime.timestamp() * 1000)
            if e.event == "view":
                views.append(ViewEvent(e.entityId, e.targetEntityId, t))
            else:
                likes.append(LikeEvent(e.entityId, e.targetEntityId, t, e.event == "like"))
        return TrainingData(users, items, views, likes)


class Preparator(PPreparator):
    def prepare(self, sc, td: TrainingData) -> PreparedData:
        return td


@dataclass
class ALSAlgorithmParams(Params):
    rank: int
    numIterations: int
    lambda_: float = field(default=0.01, metadata={"json": "lambda"})
    seed: Optional[int] = None


def _model_path(id: str) -> Path:
    return Path(os.environ.get("PIO_MODELDATA_DIR", "pio_modeldata")) / id


class ALSModel(PersistentModel):
    """productFeatures (device resident) + itemStringIntMap + items (ALSAlgorithm.scala:39-55)."""

    def __init__(self, mf: MatrixFactorizationModel, itemStringIntMap: BiMap, items: Dict[int, Item]):
        self.mf, self.itemStringIntMap, self.items = mf, itemStringIntMap, items
        self.itemIntStringMap = itemStringIntMap.inverse

    @property
    def productFeatures(self) -> Dict[int, np.ndarray]:
        return {int(i): self.mf.productFeatures[i].astype(np.float64) for i in np.flatnonzero(self.mf.productHas)}

    def save(self, id, params, sc) -> bool:
        d = _model_path(id)
        d.mkdir(parents=True, exist_ok=True)
        self.mf.save(str(d / "factors.pioals"))
        (d / "itemStringIntMap.json").write_text(json.dumps(self.itemStringIntMap.toMap()))
        (d / "items.json").write_text(json.dumps({str(k): v.categories for k, v in self.items.items()}))
        return True

    @classmethod
    def apply(cls, id, params, sc) -> "ALSModel":
        d = _model_path(id)
        mf = MatrixFactorizationModel.load(str(d / "factors.pioals"), getattr(sc, "device", 0))
        items = {int(k): Item(v) for k, v in json.loads((d / "items.json").read_text()).items()}
        return cls(mf, BiMap(json.loads((d / "itemStringIntMap.json").read_text())), items)


def candidate_mask(n_items: int, items: Dict[int, Item], queryList: Set[int], query: Query,
                   itemStringIntMap: BiMap) -> np.ndarray:
    """isCandidateItem (ALSAlgorithm.scala:237-263) as an exclusion mask (1 = not a candidate); the query items
    themselves are excluded inside the kernel."""
    mask = np.zeros(n_items, np.uint8)
    if query.whiteList is not None:
        wl = {itemStringIntMap.get(x) for x in query.whiteList}
        mask[:] = 1
        idx = [w for w in wl if w is not None]
        if idx:
            mask[idx] = 0
    if query.blackList is not None:
        idx = [b for b in (itemStringIntMap.get(x) for x in query.blackList) if b is not None]
        if idx:
            mask[idx] = 1
    if query.categories is not None or query.categoryBlackList is not None:
        for i in range(n_items):
            cats = items[i].categories if i in items else None
            if query.categories is not None:
                if cats is None or not (set(cats) & set(query.categories)):
                    mask[i] = 1
            if query.categoryBlackList is not None and cats is not None and (set(cats) & set(query.categoryBlackList)):
                mask[i] = 1
    return mask


class ALSAlgorithm(P2LAlgorithm):
    def __init__(self, ap: ALSAlgorithmParams):
        self.ap = ap

    def _ratings(self, data: PreparedData, userMap: BiMap, itemMap: BiMap):
        """view events -> ((u,i),1), unknown ids dropped; the reduceByKey(_ + _) runs on the GPU (dedup=sum)."""
        us, its = [], []
        for r in data.viewEvents:
            u, i = userMap.getOrElse(r.user, -1), itemMap.getOrElse(r.item, -1)
            if u != -1 and i != -1:
                us.append(u)
                its.append(i)
        return np.array(us, np.int32), np.array(its, np.int32), np.ones(len(us), np.float32)

    def train(self, sc, data: PreparedData) -> ALSModel:
        for name, coll in (("viewEvents", data.viewEvents), ("users", data.users), ("items", data.items)):
            if not coll:
                raise ValueError(f"requirement failed: {name} in PreparedData cannot be empty. Please check if "
                                 "DataSource generates TrainingData and Preprator generates PreparedData correctly.")
        userMap = BiMap.stringInt(data.users.keys())
        itemMap = BiMap.stringInt(data.items.keys())
        items = {itemMap(k): v for k, v in data.items.items()}
        u, i, v = self._ratings(data, userMap, itemMap)
        if u.size == 0:
            raise ValueError("requirement failed: mllibRatings cannot be empty. Please check if your events contain "
                             "valid user and item ID.")
        seed = sc.agree_seed(self.ap.seed) if hasattr(sc, "agree_seed") else (self.ap.seed or 0)
        m = ALS.trainImplicit((u, i, v), rank=self.ap.rank, iterations=self.ap.numIterations, lambda_=self.ap.lambda_,
                              blocks=-1, alpha=1.0, seed=seed, dedup=self._dedup(), n_users=userMap.size,
                              n_products=itemMap.size, sc=sc)
        return ALSModel(m, itemMap, items)

    def _dedup(self):
        return "sum"

    def predict(self, model: ALSModel, query: Query) -> PredictedResult:
        queryList = {model.itemStringIntMap.get(x) for x in query.items}
        queryList.discard(None)
        if not queryList:
            return PredictedResult([])
        mask = candidate_mask(len(model.mf.productHas), model.items, queryList, query, model.itemStringIntMap)
        items, scores, cnt = model.mf.similarProducts(sorted(queryList), query.num, mask)
        return PredictedResult([ItemScore(model.itemIntStringMap(int(items[t])), float(scores[t])) for t in range(cnt)])


class LikeAlgorithm(ALSAlgorithm):
    """like -> +1, dislike -> -1, the latest event of a (user,item) pair wins (LikeAlgorithm.scala:59-108);
    the latest-wins reduceByKey runs on the GPU (dedup=keep_last with the event times)."""

    def _ratings(self, data, userMap, itemMap):
        us, its, vs, ts = [], [], [], []
        for r in data.likeEvents:
            u, i = userMap.getOrElse(r.user, -1), itemMap.getOrElse(r.item, -1)
            if u != -1 and i != -1:
                us.append(u)
                its.append(i)
                vs.append(1.0 if r.like else -1.0)
                ts.append(r.t)
        return np.array(us, np.int32), np.array(its, np.int32), np.array(vs, np.float32), np.array(ts, np.int64)

    def train(self, sc, data):
        if not data.likeEvents:
            raise ValueError("requirement failed: likeEvents in PreparedData cannot be empty.")
        userMap = BiMap.stringInt(data.users.keys())
        itemMap = BiMap.stringInt(data.items.keys())
        items = {itemMap(k): v for k, v in data.items.items()}
        u, i, v, ts = self._ratings(data, userMap, itemMap)
        if u.size == 0:
            raise ValueError("requirement failed: mllibRatings cannot be empty.")
        seed = sc.agree_seed(self.ap.seed) if hasattr(sc, "agree_seed") else (self.ap.seed or 0)
        m = ALS.trainImplicit((u, i, v, ts), rank=self.ap.rank, iterations=self.ap.numIterations,
                              lambda_=self.ap.lambda_, blocks=-1, alpha=1.0, seed=seed, dedup="keep_last",
                              n_users=userMap.size, n_products=itemMap.size, sc=sc)
        return ALSModel(m, itemMap, items)


@dataclass
class CooccurrenceAlgorithmParams(Params):
    n: int   # top co-occurring items kept per item


class CooccurrenceModel:
    """topCooccurrences: item index -> [(other item index, count), ...] (CooccurrenceAlgorithm.scala:31-42); a plain local
    model (P2LAlgorithm keeps it as is)."""

    def __init__(self, top_items: np.ndarray, top_counts: np.ndarray, top_n: np.ndarray, itemStringIntMap: BiMap,
                 items: Dict[int, Item]):
        self.top_items, self.top_counts, self.top_n = top_items, top_counts, top_n
        self.itemStringIntMap, self.items = itemStringIntMap, items
        self.itemIntStringMap = itemStringIntMap.inverse

    def topCooccurrences(self, i: int):
        return [(int(self.top_items[i, t]), int(self.top_counts[i, t])) for t in range(int(self.top_n[i]))]


class CooccurrenceAlgorithm(P2LAlgorithm):
    """CooccurrenceAlgorithm.scala:44-175: the counting runs on the GPU (pio_cooc_train), predict sums the stored counts of
    the query items, filters and takes the top `num`."""

    def __init__(self, ap: CooccurrenceAlgorithmParams):
        self.ap = ap

    def train(self, sc, data) -> CooccurrenceModel:
        from .. import native
        itemMap = BiMap.stringInt(data.items.keys())
        userMap = BiMap.stringInt(e.user for e in data.viewEvents)     # only to index the users of the view events
        us, its = [], []
        for e in data.viewEvents:
            i = itemMap.getOrElse(e.item, -1)
            if i != -1:
                us.append(userMap(e.user))
                its.append(i)
        items = {itemMap(k): v for k, v in data.items.items()}
        n_items = itemMap.size
        if us:
            ti, tc, tn = native.cooc_train(np.array(us, np.int32), np.array(its, np.int32), max(userMap.size, 1), n_items,
                                           self.ap.n, device=getattr(sc, "device", 0))
        else:
            ti = np.full((n_items, self.ap.n), -1, np.int32)
            tc = np.zeros((n_items, self.ap.n), np.int32)
            tn = np.zeros(n_items, np.int32)
        return CooccurrenceModel(ti, tc, tn, itemMap, items)

    def predict(self, model: CooccurrenceModel, query: Query) -> PredictedResult:
        queryList = {i for i in (model.itemStringIntMap.get(x) for x in query.items) if i is not None}
        white = None if query.whiteList is None else {i for i in (model.itemStringIntMap.get(x) for x in query.whiteList)
                                                      if i is not None}
        black = None if query.blackList is None else {i for i in (model.itemStringIntMap.get(x) for x in query.blackList)
                                                      if i is not None}
        counts: Dict[int, int] = {}
        for q in queryList:
            for idx, c in model.topCooccurrences(q):
                counts[idx] = counts.get(idx, 0) + c

        def candidate(i: int) -> bool:
            if white is not None and i not in white:
                return False
            if black is not None and i in black:
                return False
            if i in queryList:
                return False
            if query.categories is not None:
                cats = model.items[i].categories if i in model.items else None
                return cats is not None and bool(set(cats) & set(query.categories))
            return True

        top = sorted(((i, v) for i, v in counts.items() if candidate(i)), key=lambda kv: (-kv[1], kv[0]))[:query.num]
        return PredictedResult([ItemScore(model.itemIntStringMap(i), float(v)) for i, v in top])


class Serving(LServing):
    """z-score standardisation per algorithm, then sum per item, top num (Serving.scala:29-69)."""

    def serve(self, query: Query, predictedResults) -> PredictedResult:
        std = []
        for pr in predictedResults:
            if query.num == 1:            # "if query 1 item, don't standardize" (Serving.scala:33-35)
                std.append(list(pr.itemScores))
                continue
            # otherwise every algorithm's scores are z-scored, also when only one algorithm is deployed (:36-57)
            sc = np.array([x.score for x in pr.itemScores], np.float64)
            mean = sc.mean() if sc.size else 0.0
            sd = sc.std(ddof=1) if sc.size > 1 else 0.0   # breeze meanAndVariance: sample variance, 0 for n <= 1
            std.append([ItemScore(x.item, 0.0 if sd == 0 else (x.score - mean) / sd) for x in pr.itemScores])
        comb: Dict[str, float] = {}
        for lst in std:
            for x in lst:
                comb[x.item] = comb.get(x.item, 0.0) + x.score
        top = sorted(comb.items(), key=lambda kv: -kv[1])[:query.num]
        return PredictedResult([ItemScore(k, v) for k, v in top])


class SimilarProductEngine(EngineFactory):
    def apply(self) -> Engine:
        return Engine(DataSource, Preparator, {"als": ALSAlgorithm, "cooccurrence": CooccurrenceAlgorithm,
                                                "likealgo": LikeAlgorithm}, Serving)
