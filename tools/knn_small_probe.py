import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "locations-recommender_b200"))
import numpy as np
import vrec
from vrec import synth
P, K, nt, unk, kern = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
inp = synth.random_knn_inputs(P, 60, 9, seed=1, separate_ratings=False)
ctx = vrec.Context(0)
rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
rs.set_option("knn_kernel", kern)
t = inp.person_id[:nt]
if unk: t = np.concatenate([t, [1, 999999]])
out = vrec.KnnRecommender(rs, 0.5, 0.5, K).recommend(t, np.arange(0, 60, 2), 10)
print("ok", P, K, nt, unk, kern, out[2][:5].tolist(), flush=True)
