"""Metric classes of the evaluation loop against the reference's own known-answer cases
(core/src/test/scala/org/apache/predictionio/controller/MetricTest.scala:63-146) and MetricEvaluator's choice of the best
engine-params set (MetricEvaluator.scala:218-262)."""
from pio_b200 import evaluation as ev


class QAverage(ev.AverageMetric):
    def calculate_one(self, q, p, a):
        return float(q)


class QOptionAverage(ev.OptionAverageMetric):
    def calculate_one(self, q, p, a):
        return None if q < 0 else float(q)


class QStdev(ev.StdevMetric):
    def calculate_one(self, q, p, a):
        return float(q)


class QOptionStdev(ev.OptionStdevMetric):
    def calculate_one(self, q, p, a):
        return None if q < 0 else float(q)


class QSum(ev.SumMetric):
    def calculate_one(self, q, p, a):
        return q


def _ds(*folds):
    return [(None, [(q, 0, 0) for q in f]) for f in folds]


def test_metric_known_answers():
    assert QAverage().calculate(None, _ds([1, 2, 3], [4, 5, 6])) == 21.0 / 6                 # "Average Metric"
    assert QOptionAverage().calculate(None, _ds([1, 2, 3], [-4, -5, 6])) == 12.0 / 4         # "Option Average Metric"
    assert QStdev().calculate(None, _ds([1, 1, 1, 1], [5, 5, 5, 5])) == 2.0                  # "Stdev Metric"
    assert QOptionStdev().calculate(None, _ds([1, 1, 1, 1], [5, 5, 5, 5, -5])) == 2.0        # "Option Stdev Metric"
    assert QSum().calculate(None, _ds([1, 2, 3], [4, 5, 6])) == 21                           # "Sum Metric [Int]"
    assert isinstance(QSum().calculate(None, _ds([1, 2, 3], [4, 5, 6])), int)
    assert ev.ZeroMetric().calculate(None, _ds([1])) == 0.0


def test_metric_evaluator_picks_the_first_maximum():
    me = ev.MetricEvaluator(QAverage(), [QSum()])
    r = me.evaluateBase(None, [("a", _ds([1, 2])), ("b", _ds([5, 5])), ("c", _ds([5, 5])), ("d", _ds([0]))])
    assert r.bestIdx == 1 and r.bestEngineParams == "b" and r.bestScore.score == 5.0 and r.bestScore.otherScores == [10]
    assert [s.score for _, s in r.engineParamsScores] == [1.5, 5.0, 5.0, 0.0]
    assert r.metricHeader == "QAverage" and r.otherMetricHeaders == ["QSum"]
