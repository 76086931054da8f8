"""`native-als`: the module that replaces the Spark-MLlib calls inside the templates.

Same names and argument meaning as the MLlib API the templates call (SURVEY 8(b) "Inner"):
    ALS.train(ratings, rank, iterations, lambda, blocks, seed)
    ALS.trainImplicit(ratings, rank, iterations, lambda, blocks, alpha, seed)
    new ALS().setRank(..).setIterations(..).setLambda(..).setImplicitPrefs(..).setAlpha(..).setSeed(..).run(ratings)
        (examples/scala-parallel-recommendation/blacklist-items/src/main/scala/ALSAlgorithm.scala:76-86)
    MatrixFactorizationModel(rank, userFeatures, productFeatures): recommendProducts, predict
    NaiveBayes.train(labeledPoints, lambda) / NaiveBayesModel.predict
        (examples/scala-parallel-classification/add-algorithm/src/main/scala/NaiveBayesAlgorithm.scala:41-57)

Ratings are COO arrays (user:int32, product:int32, rating:float32) -- the RDD[Rating(Int,Int,Double)]
the templates build at ALSAlgorithm.scala:62-65 -- and everything below is one call through the C ABI
(native.py -> libpio_als.so). No CPU path exists here.

Initial factors: MLlib seeds per-block XORShift streams whose layout depends on the executor count
(hard part 6), so `seed` selects the counter-hash initialisation (PIO_ALS_INIT_HASH)
unless explicit `init` factors are passed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import native

DEDUP = {"none": native.DEDUP_NONE, "sum": native.DEDUP_SUM, "keep_last": native.DEDUP_KEEP_LAST}


@dataclass
class Rating:
    user: int
    product: int
    rating: float


def _coo(ratings):
    if isinstance(ratings, tuple) and len(ratings) >= 3:
        u, p, r = ratings[:3]
        ts = ratings[3] if len(ratings) > 3 else None
    else:
        rs = list(ratings)
        u = np.fromiter((x.user for x in rs), np.int32, len(rs))
        p = np.fromiter((x.product for x in rs), np.int32, len(rs))
        r = np.fromiter((x.rating for x in rs), np.float32, len(rs))
        ts = None
    return (np.ascontiguousarray(u, np.int32), np.ascontiguousarray(p, np.int32),
            np.ascontiguousarray(r, np.float32), ts)


class MatrixFactorizationModel:
    """rank + factor matrices (dense, indexed by Int id; `userHas/productHas` mark ids that own a factor --
    MLlib's RDD[(Int, Array[Double])] simply lacks the others)."""

    def __init__(self, rank: int, userFeatures: np.ndarray, productFeatures: np.ndarray, userHas: np.ndarray,
                 productHas: np.ndarray, handle: Optional[native.NativeALS] = None):
        self.rank = int(rank)
        self.userFeatures = userFeatures
        self.productFeatures = productFeatures
        self.userHas = userHas
        self.productHas = productHas
        self._h = handle

    def _handle(self, device: int = 0) -> native.NativeALS:
        if self._h is None:
            raise native.NativeError(native.ERR_STATE, "model has no device handle; load it with ALS.load / from file")
        return self._h

    def recommendProducts(self, user: int, num: int) -> list:
        items, scores, cnt = self._handle().recommend(np.array([user], np.int32), num)
        return [Rating(user, int(items[0, t]), float(scores[0, t])) for t in range(int(cnt[0]))]

    def recommendProductsWithFilter(self, user: int, num: int, productIdFilter: Sequence[int]) -> list:
        mask = np.zeros(self.productFeatures.shape[0], np.uint8)
        idx = np.fromiter((int(i) for i in productIdFilter), np.int64)
        if idx.size:
            mask[idx] = 1
        items, scores, cnt = self._handle().recommend(np.array([user], np.int32), num, mask)
        return [Rating(user, int(items[0, t]), float(scores[0, t])) for t in range(int(cnt[0]))]

    def recommendProductsForUsers(self, users: np.ndarray, num: int, item_mask: Optional[np.ndarray] = None,
                                  item_weight: Optional[np.ndarray] = None):
        """Batched top-N (what batchPredict's cartesian + groupBy computes, ALSAlgorithm.scala:117-158)."""
        return self._handle().recommend(np.ascontiguousarray(users, np.int32), num, item_mask, item_weight)

    def similarProducts(self, query_items: Sequence[int], num: int, item_mask: Optional[np.ndarray] = None,
                        item_weight: Optional[np.ndarray] = None, exclude_query: bool = True):
        """exclude_query=True: similarproduct's `!queryList.contains(i)` rule; False: ecommerce predictSimilar, whose
        isCandidateItem has no such rule (train-with-rate-event ECommAlgorithm.scala:492-525,527-557)."""
        return self._handle().similar(np.asarray(list(query_items), np.int32), num, item_mask, item_weight,
                                      keep_query_items=not exclude_query)

    def predict(self, user: int, product: int) -> float:
        return float(np.dot(self.userFeatures[user].astype(np.float64), self.productFeatures[product].astype(np.float64)))

    def save(self, path: str) -> None:
        self._handle().save(path)

    @staticmethod
    def load(path: str, device: int = 0) -> "MatrixFactorizationModel":
        h = native.NativeALS.load(path, device)
        uf, pf, uh, ph = h.get_factors()
        return MatrixFactorizationModel(h.rank, uf, pf, uh, ph, h)


class ALS:
    def __init__(self):
        self.rank, self.iterations, self.lambda_, self.implicitPrefs = 10, 10, 0.01, False
        self.alpha, self.seed, self.userBlocks, self.productBlocks, self.checkpointInterval = 1.0, 0, -1, -1, 10
        self.dedup, self.device = "none", 0

    # builder (mllib.recommendation.ALS setters)
    def setRank(self, v): self.rank = int(v); return self
    def setIterations(self, v): self.iterations = int(v); return self
    def setLambda(self, v): self.lambda_ = float(v); return self
    def setImplicitPrefs(self, v): self.implicitPrefs = bool(v); return self
    def setAlpha(self, v): self.alpha = float(v); return self
    def setSeed(self, v): self.seed = int(v); return self
    def setUserBlocks(self, v): self.userBlocks = int(v); return self        # accepted, meaningless on one GPU
    def setProductBlocks(self, v): self.productBlocks = int(v); return self
    def setCheckpointInterval(self, v): self.checkpointInterval = int(v); return self
    def setDedup(self, mode): self.dedup = mode; return self                 # extension: GPU-side reduceByKey
    def setDevice(self, d): self.device = int(d); return self

    def run(self, ratings, n_users: Optional[int] = None, n_products: Optional[int] = None,
            init: Optional[Tuple[np.ndarray, Optional[np.ndarray]]] = None, sc=None) -> MatrixFactorizationModel:
        u, p, r, ts = _coo(ratings)
        if u.shape[0] == 0:
            raise ValueError("requirement failed: ratings cannot be empty")
        nu = int(n_users) if n_users is not None else int(u.max()) + 1
        npr = int(n_products) if n_products is not None else int(p.max()) + 1
        world, wrank, nccl_id = 1, 0, None
        device = self.device
        if sc is not None:
            device = getattr(sc, "device", device)
            world, wrank, nccl_id = getattr(sc, "world_size", 1), getattr(sc, "world_rank", 0), None
            if world > 1:
                nccl_id = sc.new_nccl_id()
        h = native.NativeALS(self.rank, nu, npr, lam=self.lambda_, implicit=self.implicitPrefs, alpha=self.alpha,
                             seed=self.seed, device=device, world_size=world, world_rank=wrank, nccl_id=nccl_id,
                             init_mode=native.INIT_CALLER if init is not None else native.INIT_HASH)
        uf, pf, uh, ph = h.train(u, p, r, self.iterations, dedup=DEDUP[self.dedup], ts=ts,
                                 user_init=None if init is None else init[0],
                                 item_init=None if init is None else init[1])
        return MatrixFactorizationModel(self.rank, uf, pf, uh, ph, h)

    @staticmethod
    def train(ratings, rank, iterations, lambda_=0.01, blocks=-1, seed=0, **kw) -> MatrixFactorizationModel:
        return ALS().setRank(rank).setIterations(iterations).setLambda(lambda_).setSeed(seed) \
            .setDedup(kw.pop("dedup", "none")).run(ratings, **kw)

    @staticmethod
    def trainImplicit(ratings, rank, iterations, lambda_=0.01, blocks=-1, alpha=1.0, seed=0, **kw) -> MatrixFactorizationModel:
        return ALS().setRank(rank).setIterations(iterations).setLambda(lambda_).setImplicitPrefs(True) \
            .setAlpha(alpha).setSeed(seed).setDedup(kw.pop("dedup", "none")).run(ratings, **kw)


class NaiveBayesModel:
    def __init__(self, labels: np.ndarray, pi: np.ndarray, theta: np.ndarray, device: int = 0):
        self.labels, self.pi, self.theta, self.device = labels, pi, theta, device

    def predict(self, features) -> float:
        x = np.asarray(features, np.float32).reshape(1, -1)
        return float(self.labels[native.nb_predict(x, self.pi, self.theta, self.device)[0]])

    def predictBatch(self, x: np.ndarray) -> np.ndarray:
        return self.labels[native.nb_predict(np.asarray(x, np.float32), self.pi, self.theta, self.device)]


class NaiveBayes:
    @staticmethod
    def train(labels: np.ndarray, features: np.ndarray, lambda_: float = 1.0, device: int = 0) -> NaiveBayesModel:
        """labels: float label values (LabeledPoint.label), features: n x F non-negative."""
        labels = np.asarray(labels, np.float64)
        x = np.asarray(features, np.float32)
        if (x < 0).any():
            raise ValueError("Naive Bayes requires nonnegative feature values")  # MLlib requirement
        classes, idx = np.unique(labels, return_inverse=True)                     # sorted ascending, as MLlib
        pi, theta = native.nb_train(idx.astype(np.int32), x, classes.shape[0], lambda_, device)
        return NaiveBayesModel(classes, pi, theta, device)
