package com.github.tashoyan.recommender.stochastic

import com.github.tashoyan.recommender.gpu.{GpuContext, Vrec}
import org.apache.spark.sql.{DataFrame, SparkSession}

/**
  * Drop-in for recommender/src/main/scala/com/github/tashoyan/recommender/stochastic/StochasticRecommender.scala:
  * same constructor and `require`s (:28-34), same `makeRecommendations(vertexId): DataFrame(id, probability)`
  * with `id != vertexId and probability > 0` (:66-90), the same println of step() (:94,:100) and the same
  * IllegalArgumentException for an unknown vertex (:70).  The edge DataFrame is collected once; the damped power
  * iteration with its convergence test runs on the device (no Spark action per iteration).
  */
class StochasticRecommender(
  stochasticEdges: DataFrame,
  epsilon: Double,
  maxIterations: Int
)(implicit spark: SparkSession) extends AutoCloseable {
  require(epsilon >= 0, "epsilon must be non-negative")
  require(maxIterations >= 0, "max iterations number must be non-negative")

  private lazy val handle: Long = {
    val edges = stochasticEdges.select("source_id", "target_id", "balanced_weight").collect()
    val h = Vrec.sgLoad(GpuContext.ctx, edges.map(_.getLong(0)), edges.map(_.getLong(1)), edges.map(_.getDouble(2)))
    if (h <= 0) throw new IllegalStateException(s"vrec_sg_load failed: ${Vrec.lastError()}")
    h
  }
  private lazy val vertexIds: Array[Long] = {
    val ids = new Array[Long](Vrec.sgVertexCount(handle).toInt)
    if (Vrec.sgVertexIds(handle, ids) != Vrec.OK) throw new IllegalStateException(Vrec.lastError())
    ids
  }

  private def printStep(iterations: Int, converged: Int): Unit =
    if (converged == 1) println(s"Converged in $iterations iterations")
    else println(s"Number of iterations $iterations reached the maximum $maxIterations")

  def makeRecommendations(vertexId: Long): DataFrame = {
    import spark.implicits._
    val x = new Array[Double](vertexIds.length)
    val info = new Array[Int](2)
    val rc = Vrec.sgStationary(handle, vertexId, epsilon, maxIterations, x, info)
    if (rc == Vrec.ENOENT) throw new IllegalArgumentException(s"No such vertex in the graph: $vertexId")
    if (rc != Vrec.OK) throw new IllegalStateException(Vrec.lastError())
    printStep(info(0), info(1))
    vertexIds.zip(x).filter { case (id, p) => id != vertexId && p > 0 }.toSeq.toDF("id", "probability")
  }

  /**
    * makeRecommendations + StochasticRecommenderMain.printRecommendations' join with the target region's places
    * and `orderBy(probability desc).limit(maxRecommendations)` (StochasticRecommenderMain.scala:69-73).
    */
  def recommend(vertexId: Long, regionPlaceIds: Array[Long], maxRecommendations: Int): Seq[(Long, Double)] = {
    val m = math.max(1, maxRecommendations)
    val outId = new Array[Long](m)
    val outProb = new Array[Double](m)
    val one = Array.fill(4)(new Array[Int](1))
    val rc = Vrec.sgQuery(handle, Array(vertexId), epsilon, maxIterations, regionPlaceIds, maxRecommendations, outId,
      outProb, one(0), one(1), one(2), one(3))
    if (rc != Vrec.OK) throw new IllegalStateException(Vrec.lastError())
    if (one(3)(0) == Vrec.ENOENT) throw new IllegalArgumentException(s"No such vertex in the graph: $vertexId")
    printStep(one(1)(0), one(2)(0))
    outId.take(one(0)(0)).zip(outProb.take(one(0)(0))).toSeq
  }

  override def close(): Unit = Vrec.sgFree(handle)
}
