"""vrec -- host-side mirror of the reference's recommender interface over libvrec.so.

`KnnRecommender` and `StochasticRecommender` keep the constructor arguments, method names and
error behaviour of the Scala classes (knn/KnnRecommender.scala:9-25,
stochastic/StochasticRecommender.scala:28-71); the compute runs in hand-written sm_100a CUDA
kernels behind the C ABI of include/vrec.h.
"""
from .engine import (Context, KnnRecommender, KnnRegionSet, StochasticGraph,  # noqa: F401
                     StochasticGraphGroup, StochasticRecommender, VrecError, NoSuchElement)
