package com.github.tashoyan.recommender.knn

import com.github.tashoyan.recommender.gpu.{GpuContext, Vrec}
import org.apache.spark.ml.linalg.SparseVector
import org.apache.spark.sql.functions.col
import org.apache.spark.sql.{DataFrame, SparkSession}

/**
  * Drop-in for recommender/src/main/scala/com/github/tashoyan/recommender/knn/KnnRecommender.scala:
  * same constructor, same `require`s (:17-20), same `makeRecommendations(personId): DataFrame(place_id,
  * estimated_rating)` (:22-25) and the same IllegalArgumentException for an unknown person (:83).
  * The three DataFrames are collected once into the CSR arrays vrec_knn_load takes; every query then runs on the
  * device.  `KnnRecommenderMain.makeRecommendations` (KnnRecommenderMain.scala:53-67) builds a recommender per
  * query: keep one instance per region-set (`Seq(home, target).sorted.distinct`) to keep the data resident.
  */
class KnnRecommender(
  placeRatingVectors: DataFrame,
  categoryRatingVectors: DataFrame,
  placeRatings: DataFrame,
  placeWeight: Double,
  categoryWeight: Double,
  kNearest: Int
) extends AutoCloseable {
  private val spark: SparkSession = placeRatings.sparkSession // the reference's constructor takes no session (:9-16)

  require(placeWeight > 0 && placeWeight < 1.0, s"Place weight must be in the interval (0; 1): $placeWeight")
  require(categoryWeight > 0 && categoryWeight < 1.0, s"Category weight must be in the interval (0; 1): $categoryWeight")
  require(placeWeight + categoryWeight == 1.0, s"Sum of weights must be 1.0: place: $placeWeight, category: $categoryWeight")
  require(kNearest > 0, "K nearest must be positive")

  private case class Csr(persons: Array[Long], rowPtr: Array[Long], cols: Array[Int], vals: Array[Double], dim: Int)

  /** (person_id, rating_vector: SparseVector) rows -> CSR over `persons` (a person without a row gets an empty one). */
  private def toCsr(vectors: DataFrame, persons: Array[Long]): Csr = {
    val rows = vectors
      .select(col("person_id"), col("rating_vector"))
      .collect()
      .map(r => r.getLong(0) -> r.getAs[SparseVector](1))
      .toMap
    val rowPtr = new Array[Long](persons.length + 1)
    val cols = Array.newBuilder[Int]
    val vals = Array.newBuilder[Double]
    var dim = 1
    var i = 0
    while (i < persons.length) {
      rows.get(persons(i)).foreach { v =>
        cols ++= v.indices
        vals ++= v.values
        dim = math.max(dim, v.size)
        rowPtr(i + 1) = v.indices.length.toLong
      }
      i += 1
    }
    i = 0
    while (i < persons.length) { rowPtr(i + 1) += rowPtr(i); i += 1 }
    Csr(persons, rowPtr, cols.result(), vals.result(), dim)
  }

  private lazy val handle: Long = {
    val persons = placeRatingVectors.select("person_id")
      .union(categoryRatingVectors.select("person_id"))
      .distinct()
      .collect()
      .map(_.getLong(0))
      .sorted
    val p = toCsr(placeRatingVectors, persons)
    val c = toCsr(categoryRatingVectors, persons)
    val ratings = placeRatings.select("person_id", "place_id", "rating").collect()
    val h = Vrec.knnLoad(
      GpuContext.ctx, persons, p.rowPtr, p.cols, p.vals, p.dim, c.rowPtr, c.cols, c.vals, c.dim,
      ratings.map(_.getLong(0)), ratings.map(_.getLong(1)), ratings.map(_.getLong(2))
    )
    if (h <= 0) throw new IllegalStateException(s"vrec_knn_load failed: ${Vrec.lastError()}")
    h
  }
  private lazy val placeDim: Int = math.max(1 << 16, placeRatings.select("place_id").distinct().count().toInt + 1)

  def makeRecommendations(personId: Long): DataFrame = {
    import spark.implicits._
    var capacity = placeDim
    while (true) {
      val outPlace = new Array[Long](capacity)
      val outRating = new Array[Double](capacity)
      val n = Vrec.knnEstimates(handle, personId, placeWeight, categoryWeight, kNearest, outPlace, outRating)
      if (n == Vrec.ENOENT) throw new IllegalArgumentException(s"No such person: $personId")
      if (n < 0) throw new IllegalStateException(Vrec.lastError())
      if (n < capacity) {
        return outPlace.take(n.toInt).zip(outRating.take(n.toInt)).toSeq.toDF("place_id", "estimated_rating")
      }
      capacity *= 4
    }
    throw new IllegalStateException("unreachable")
  }

  /**
    * makeRecommendations + KnnRecommenderMain.printRecommendations' region filter and
    * `orderBy(estimated_rating desc).limit(maxRecommendations)` (KnnRecommenderMain.scala:96-102) in one device
    * pass: (place_id, estimated_rating) in that order, ties by place_id.
    */
  def recommend(personId: Long, regionPlaceIds: Array[Long], maxRecommendations: Int): Seq[(Long, Double)] = {
    val m = math.max(1, maxRecommendations)
    val outPlace = new Array[Long](m)
    val outRating = new Array[Double](m)
    val outCount = new Array[Int](1)
    val outStatus = new Array[Int](1)
    val rc = Vrec.knnQuery(handle, Array(personId), placeWeight, categoryWeight, kNearest, regionPlaceIds,
      maxRecommendations, outPlace, outRating, outCount, outStatus)
    if (rc != Vrec.OK) throw new IllegalStateException(Vrec.lastError())
    if (outStatus(0) == Vrec.ENOENT) throw new IllegalArgumentException(s"No such person: $personId")
    outPlace.take(outCount(0)).zip(outRating.take(outCount(0))).toSeq
  }

  override def close(): Unit = Vrec.knnFree(handle)
}
