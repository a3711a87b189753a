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

# ---------------------------------------------------------------- builders
# LocationTest.scala:8-29: same location -> 0; two Moscow locations -> 745 +- 5 m; commutative
LOCATION_SAME = (55.6438965, 37.4433515)
LOCATION_PAIR = ((55.612652, 37.591753), (55.611152, 37.603366))
LOCATION_PAIR_DISTANCE, LOCATION_PAIR_TOLERANCE = 745.0, 5.0

# StochasticGraphBuilderTest.scala:11-66: five edge families with weights that sum to 1 per source, betas
# (0.4, 0.6, 1.0, 0.3, 0.7); after buildWithBalancedWeights the balanced weights of every source sum to 1.0
# (`===`, exact).  Families as (beta, [(source, target, weight)]); the weights 0.4 / 0.6 are 2/5 and 3/5.
SGB_FAMILIES = [
    (0.4, [(1, 2, 1.0)]),
    (0.6, [(1, 3, 0.4), (1, 5, 0.6)]),
    (1.0, [(3, 5, 1.0), (5, 3, 1.0)]),
    (0.3, [(2, 4, 1.0), (4, 2, 1.0)]),
    (0.7, [(2, 3, 1.0), (4, 5, 1.0)]),
]
