/*
 * The recommendation template's model on top of native-als.  Patch of
 * examples/scala-parallel-recommendation/blacklist-items/src/main/scala/ALSModel.scala: the model owns device memory,
 * so -- exactly like the reference's RDD-backed model -- it is a PersistentModel (controller/PersistentModel.scala:67-103):
 * save() writes one factor file (pio_als_save) plus the two BiMaps, the companion's apply() reloads them on deploy.
 */
package org.example.recommendation

import org.apache.predictionio.controller.PersistentModel
import org.apache.predictionio.controller.PersistentModelLoader
import org.apache.predictionio.data.storage.BiMap
import org.apache.predictionio.nativeals.NativeModel

import org.apache.spark.SparkContext

class ALSModel(
    val native: NativeModel,
    val userStringIntMap: BiMap[String, Int],
    val itemStringIntMap: BiMap[String, Int])
  extends PersistentModel[ALSAlgorithmParams] {

  def save(id: String, params: ALSAlgorithmParams, sc: SparkContext): Boolean = {
    new java.io.File(s"/tmp/${id}").mkdirs()
    native.save(s"/tmp/${id}/factors.pioals")
    sc.parallelize(Seq(userStringIntMap)).saveAsObjectFile(s"/tmp/${id}/userStringIntMap")
    sc.parallelize(Seq(itemStringIntMap)).saveAsObjectFile(s"/tmp/${id}/itemStringIntMap")
    true
  }

  override def toString = s"native ALS model: rank ${native.rank}, ${native.nUsers} users, ${native.nItems} items"
}

object ALSModel extends PersistentModelLoader[ALSAlgorithmParams, ALSModel] {
  def apply(id: String, params: ALSAlgorithmParams, sc: Option[SparkContext]) = {
    val users = sc.get.objectFile[BiMap[String, Int]](s"/tmp/${id}/userStringIntMap").first
    val items = sc.get.objectFile[BiMap[String, Int]](s"/tmp/${id}/itemStringIntMap").first
    new ALSModel(NativeModel.load(s"/tmp/${id}/factors.pioals", params.rank, users.size, items.size), users, items)
  }
}
