package com.github.tashoyan.recommender.gpu

/**
  * JNI binding of libvrec.so (include/vrec.h); the native side is jni/vrec_jni.c, one stub per method.
  * Handles are `Long`s (> 0); functions that return a handle or a row count return a negative VREC_E* code on
  * failure, `lastError()` then holds the message.
  */
object Vrec {
  System.loadLibrary("vrec_jni") // libvrec_jni.so links libvrec.so

  val OK: Int = 0
  val ENOENT: Int = -2 // "No such person: <id>" / "No such vertex in the graph: <id>"
  val ENOMEM: Int = -12
  val EINVAL: Int = -22 // a `require(...)` of the reference would have failed

  @native def init(device: Int): Long // vrec_init
  @native def shutdown(ctx: Long): Unit // vrec_shutdown
  @native def lastError(): String // vrec_last_error

  @native def knnLoad(
    ctx: Long,
    personId: Array[Long],
    placeRowPtr: Array[Long],
    placeCol: Array[Int],
    placeVal: Array[Double],
    placeDim: Int,
    catRowPtr: Array[Long],
    catCol: Array[Int],
    catVal: Array[Double],
    catDim: Int,
    ratingPerson: Array[Long],
    ratingPlace: Array[Long],
    ratingValue: Array[Long]
  ): Long // vrec_knn_load
  @native def knnFree(knn: Long): Unit // vrec_knn_free
  @native def knnQuery(
    knn: Long,
    targets: Array[Long],
    placeWeight: Double,
    categoryWeight: Double,
    kNearest: Int,
    placeFilter: Array[Long],
    maxRecs: Int,
    outPlace: Array[Long],
    outRating: Array[Double],
    outCount: Array[Int],
    outStatus: Array[Int]
  ): Int // vrec_knn_query
  @native def knnEstimates(
    knn: Long,
    target: Long,
    placeWeight: Double,
    categoryWeight: Double,
    kNearest: Int,
    outPlace: Array[Long],
    outRating: Array[Double]
  ): Long // vrec_knn_estimates
  @native def knnNeighbours(
    knn: Long,
    target: Long,
    placeWeight: Double,
    categoryWeight: Double,
    kNearest: Int,
    outPerson: Array[Long],
    outSimilarity: Array[Double]
  ): Int // vrec_knn_neighbours

  @native def sgLoad(ctx: Long, source: Array[Long], target: Array[Long], weight: Array[Double]): Long // vrec_sg_load
  @native def sgFree(sg: Long): Unit // vrec_sg_free
  @native def sgVertexCount(sg: Long): Long // vrec_sg_vertex_count
  @native def sgVertexIds(sg: Long, outIds: Array[Long]): Int // vrec_sg_vertex_ids
  @native def sgStationary(
    sg: Long,
    vertex: Long,
    epsilon: Double,
    maxIterations: Int,
    outX: Array[Double],
    outInfo: Array[Int]
  ): Int // vrec_sg_stationary; outInfo = (iterations, converged)
  @native def sgQuery(
    sg: Long,
    vertices: Array[Long],
    epsilon: Double,
    maxIterations: Int,
    placeFilter: Array[Long],
    maxRecs: Int,
    outId: Array[Long],
    outProb: Array[Double],
    outCount: Array[Int],
    outIterations: Array[Int],
    outConverged: Array[Int],
    outStatus: Array[Int]
  ): Int // vrec_sg_query

  // builders: PlaceVisits.calcPlaceVisits, RatingsBuilder + RatingVectorsBuilder, StochasticGraphBuilderMain
  @native def buildPlaceVisits(
    ctx: Long,
    person: Array[Long],
    lat: Array[Double],
    lon: Array[Double],
    timestampMs: Array[Long],
    region: Array[Long],
    placeId: Array[Long],
    placeLat: Array[Double],
    placeLon: Array[Double],
    placeCategory: Array[Long],
    placeRegion: Array[Long],
    lastDaysCount: Int,
    accuracyMeters: Double,
    outPerson: Array[Long],
    outTimestampMs: Array[Long],
    outPlace: Array[Long],
    outRegion: Array[Long],
    outCategory: Array[Long],
    outN: Array[Long]
  ): Long // vrec_build_place_visits
  @native def buildRatingVectors(
    ctx: Long,
    person: Array[Long],
    entity: Array[Long],
    weight: Array[Long],
    topN: Int,
    outPerson: Array[Long],
    outRowPtr: Array[Long],
    outCol: Array[Int],
    outVal: Array[Double],
    outInfo: Array[Long]
  ): Int // vrec_build_rating_vectors; outInfo = (persons, non-zeros, vector size)
  @native def buildStochasticGraph(
    ctx: Long,
    person: Array[Long],
    place: Array[Long],
    category: Array[Long],
    timestampMs: Array[Long],
    betaPersonPlace: Double,
    betaPersonCategory: Double,
    outSource: Array[Long],
    outTarget: Array[Long],
    outWeight: Array[Double],
    outN: Array[Long]
  ): Long // vrec_build_stochastic_graph
}

/** One engine context per JVM (the reference has one driver and one REPL thread: KnnRecommenderMain.scala:35-49). */
object GpuContext {
  lazy val ctx: Long = {
    val c = Vrec.init(-1)
    if (c == 0) throw new IllegalStateException(s"vrec_init failed: ${Vrec.lastError()}")
    sys.addShutdownHook(Vrec.shutdown(c))
    c
  }
}
