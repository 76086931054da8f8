"""`pio eval` on top of the GPU path: metrics, MetricEvaluator and the k-fold evaluation loop.

Mirrors core/src/main/scala/org/apache/predictionio/controller/Metric.scala (AverageMetric :99-121,
OptionAverageMetric :124-148, StdevMetric :151-176, OptionStdevMetric :179-202, SumMetric :205-231),
MetricEvaluator.scala (evaluateBase :218-262: best = the first engine-params set with the maximal primary score) and
the batch loop of Engine.eval (Engine.scala:728-817).  Every fold of every engine-params set is one ALS training and one
batched top-k call on the device (the reference's sample evaluation is 3 x 3 parameter sets x 5 folds = 45
trainings); the metric arithmetic itself is host-side bookkeeping.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple


class Metric:
    """calculate(sc, evalDataSet) with evalDataSet = [(evalInfo, [(q, p, a), ...]), ...]; larger is better."""

    @property
    def header(self) -> str:
        return type(self).__name__

    def calculate(self, sc, evalDataSet):
        raise NotImplementedError

    def compare(self, r0, r1) -> int:
        return (r0 > r1) - (r0 < r1)


def _values(metric, evalDataSet, optional: bool) -> List[float]:
    out = []
    for _, qpas in evalDataSet:
        for q, p, a in qpas:
            v = metric.calculate_one(q, p, a)
            if optional and v is None:
                continue
            out.append(float(v))
    return out


class AverageMetric(Metric):
    def calculate_one(self, q, p, a) -> float:
        raise NotImplementedError

    def calculate(self, sc, evalDataSet) -> float:
        v = _values(self, evalDataSet, False)
        return sum(v) / len(v) if v else float("nan")


class OptionAverageMetric(AverageMetric):
    """calculate_one may return None: such (q, p, a) are left out of the mean."""

    def calculate(self, sc, evalDataSet) -> float:
        v = _values(self, evalDataSet, True)
        return sum(v) / len(v) if v else float("nan")


class StdevMetric(AverageMetric):
    """Population standard deviation (StatCounter.stdev)."""

    def calculate(self, sc, evalDataSet) -> float:
        v = _values(self, evalDataSet, isinstance(self, OptionStdevMetric))
        if not v:
            return float("nan")
        m = sum(v) / len(v)
        return math.sqrt(sum((x - m) ** 2 for x in v) / len(v))


class OptionStdevMetric(StdevMetric):
    pass


class SumMetric(Metric):
    def calculate_one(self, q, p, a):
        raise NotImplementedError

    def calculate(self, sc, evalDataSet):
        tot = 0
        for _, qpas in evalDataSet:
            for q, p, a in qpas:
                tot = tot + self.calculate_one(q, p, a)
        return tot


class ZeroMetric(Metric):
    def calculate(self, sc, evalDataSet) -> float:
        return 0.0


@dataclass
class MetricScores:
    score: Any
    otherScores: List[Any]


@dataclass
class MetricEvaluatorResult:
    bestScore: MetricScores
    bestEngineParams: Any
    bestIdx: int
    metricHeader: str
    otherMetricHeaders: List[str]
    engineParamsScores: List[Tuple[Any, MetricScores]]


class MetricEvaluator:
    def __init__(self, metric: Metric, otherMetrics: Sequence[Metric] = ()):
        self.metric, self.otherMetrics = metric, list(otherMetrics)

    def evaluateBase(self, sc, engineEvalDataSet) -> MetricEvaluatorResult:
        results = [(ep, MetricScores(self.metric.calculate(sc, ds), [m.calculate(sc, ds) for m in self.otherMetrics]))
                   for ep, ds in engineEvalDataSet]
        best = 0
        for i in range(1, len(results)):   # reduce { (x, y) => if (compare(x, y) >= 0) x else y }: first maximum wins
            if self.metric.compare(results[best][1].score, results[i][1].score) < 0:
                best = i
        return MetricEvaluatorResult(results[best][1], results[best][0], best, self.metric.header,
                                     [m.header for m in self.otherMetrics], results)


class Evaluation:
    """engineEvaluator = (engine, evaluator) (controller/Evaluation.scala)."""
    engine = None
    evaluator: Optional[MetricEvaluator] = None


class EngineParamsGenerator:
    engineParamsList: List[Any] = []


def run_evaluation(evaluation: Evaluation, generator: EngineParamsGenerator, sc=None) -> MetricEvaluatorResult:
    """CoreWorkflow.runEvaluation / EvaluationWorkflow.runEvaluation: Engine.batchEval over the parameter sets, then the
    evaluator.  One WorkflowContext (= one GPU) serves all trainings."""
    from .workflow import WorkflowContext
    sc = sc or WorkflowContext(mode="Evaluation")
    engine = evaluation.engine
    data = []
    for ep in generator.engineParamsList:
        folds = engine.eval(sc, ep)                      # [(evalInfo, [(q, p, a), ...]), ...] -- one training per fold
        data.append((ep, folds))
    return evaluation.evaluator.evaluateBase(sc, data)
