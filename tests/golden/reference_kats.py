"""Known-answer tests transcribed from the reference's own test suites.

Sources (under /root/reference/recommender/src/test/scala/com/github/tashoyan/recommender/):
  knn/DistanceTest.scala:10-60
  stochastic/StochasticRecommenderTest.scala:11-21 (graph), :53-58, :76-81 (vectors), :85-93 (error)
The reference compares with exact `should be` equality; so do we.
"""
import math

# (indices, values) -> expected vectorLength      DistanceTest.scala:10-32
VECTOR_LENGTH = [
    (([], []), 0.0),
    (([0], [1.0]), 1.0),
    (([0, 1], [3.0, 4.0]), 5.0),
    (([0, 1], [-3.0, -4.0]), 5.0),
]

# ((idx1, val1), (idx2, val2)) -> expected cosineSimilarity   DistanceTest.scala:34-60
COSINE = [
    ((([0], [2.0]), ([0], [3.0])), 1.0),
    ((([0], [2.0]), ([0], [-3.0])), -1.0),
    ((([0], [2.0]), ([1], [3.0])), 0.0),
    ((([0], [2.0]), ([0, 1], [1.0, 1.0])), 1 / math.sqrt(2)),
]

# StochasticRecommenderTest.scala:11-21
SG_EDGES = [
    (1, 2, 0.4), (1, 3, 0.24), (1, 5, 0.36),
    (2, 4, 0.3), (2, 3, 0.7),
    (3, 5, 1.0),
    (4, 2, 0.3), (4, 5, 0.7),
    (5, 3, 1.0),
]

# (vertex, epsilon, maxIterations) -> [(id, probability)] sorted by -probability
SG_CASES = [
    ((1, 0.01, 1), [(5, 0.3502), (3, 0.3298), (2, 0.11900000000000001), (4, 0.051)]),
    ((1, 0.05, 1000), [(3, 0.408242766375), (5, 0.3716171248749999), (2, 0.055161925125),
                       (4, 0.014978183624999999)]),
]
SG_MISSING_VERTEX = 100  # -> IllegalArgumentException("No such vertex in the graph: 100")
