/*
 * The recommendation template's algorithm with the MLlib call replaced by native-als.
 * Patch of examples/scala-parallel-recommendation/blacklist-items/src/main/scala/ALSAlgorithm.scala:
 *   lines 52-65  (require, the two BiMap.stringInt, the index mapping)  -- unchanged
 *   lines 76-86  (new ALS()....run(mllibRatings))                        -- NativeALS.run
 *   lines 88-94  (new ALSModel(...))                                      -- the model keeps the native handle
 *   lines 97-158 (predict / batchPredict)                                 -- pio_als_recommend through NativeModel
 * Engine.scala, DataSource.scala, Preparator.scala, Serving.scala and engine.json are untouched: the class is still
 * registered as Map("als" -> classOf[ALSAlgorithm]) and constructed by Doer from engine.json's params.
 */
package org.example.recommendation

import org.apache.predictionio.controller.PAlgorithm
import org.apache.predictionio.controller.Params
import org.apache.predictionio.data.storage.BiMap
import org.apache.predictionio.nativeals.NativeALS

import org.apache.spark.SparkContext
import org.apache.spark.SparkContext._
import org.apache.spark.rdd.RDD
import org.apache.spark.mllib.recommendation.{Rating => MLlibRating}

import grizzled.slf4j.Logger

case class ALSAlgorithmParams(
  rank: Int,
  numIterations: Int,
  lambda: Double,
  seed: Option[Long]) extends Params

class ALSAlgorithm(val ap: ALSAlgorithmParams)
  extends PAlgorithm[PreparedData, ALSModel, Query, PredictedResult] {

  @transient lazy val logger = Logger[this.type]

  def train(sc: SparkContext, data: PreparedData): ALSModel = {
    // MLLib ALS cannot handle empty training data.
    require(!data.ratings.take(1).isEmpty,
      s"RDD[Rating] in PreparedData cannot be empty." +
      " Please check if DataSource generates TrainingData" +
      " and Preparator generates PreparedData correctly.")
    // Convert user and item String IDs to Int index
    val userStringIntMap = BiMap.stringInt(data.ratings.map(_.user))
    val itemStringIntMap = BiMap.stringInt(data.ratings.map(_.item))
    val mllibRatings = data.ratings.map( r =>
      // MLlibRating requires integer index for user and item
      MLlibRating(userStringIntMap(r.user), itemStringIntMap(r.item), r.rating)
    )

    // seed for the initial factors
    val seed = ap.seed.getOrElse(System.nanoTime)

    // was: new ALS().setUserBlocks(-1)....setSeed(seed).run(mllibRatings)
    val m = NativeALS.run(mllibRatings, userStringIntMap.size, itemStringIntMap.size, ap.rank, ap.numIterations,
      ap.lambda, implicitPrefs = false, alpha = 1.0, seed = seed)

    new ALSModel(m, userStringIntMap, itemStringIntMap)
  }

  def predict(model: ALSModel, query: Query): PredictedResult = {
    // Convert String ID to Int index
    model.userStringIntMap.get(query.user).map { userInt =>
      val itemIntStringMap = model.itemStringIntMap.inverse
      // blackList as a byte mask over the item indices (was: recommendProductsWithFilter's filter)
      val mask: Array[Byte] = query.blackList.map { bl =>
        val m = new Array[Byte](model.itemStringIntMap.size)
        bl.flatMap(model.itemStringIntMap.get).foreach(i => m(i) = 1)
        m
      }.orNull
      val itemScores = model.native.recommendProducts(Array(userInt), query.num, mask).head
        .map { case (i, s) => ItemScore(itemIntStringMap(i), s) }
      PredictedResult(itemScores)
    }.getOrElse {
      logger.info(s"No prediction for unknown user ${query.user}.")
      PredictedResult(Array.empty)
    }
  }

  // batch form: one native call for all queries (the reference's cartesian + groupBy, ALSAlgorithm.scala:117-158)
  override
  def batchPredict(model: ALSModel, queries: RDD[(Long, Query)]): RDD[(Long, PredictedResult)] = {
    val qs = queries.collect()
    val known = qs.flatMap { case (ix, q) => model.userStringIntMap.get(q.user).map(u => (ix, q, u)) }
    val num = if (known.isEmpty) 0 else known.map(_._2.num).max
    val itemIntStringMap = model.itemStringIntMap.inverse
    val res = if (known.isEmpty) Array.empty[Array[(Int, Double)]]
      else model.native.recommendProducts(known.map(_._3), num)
    val byIx = known.zip(res).map { case ((ix, q, _), r) =>
      ix -> PredictedResult(r.take(q.num).map { case (i, s) => ItemScore(itemIntStringMap(i), s) })
    }.toMap
    queries.sparkContext.parallelize(qs.map { case (ix, _) => (ix, byIx.getOrElse(ix, PredictedResult(Array.empty))) })
  }
}
