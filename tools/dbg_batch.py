import sys, numpy as np
sys.path.insert(0,'/root/repo')
from monoorbslam3_b200 import ORBExtractor, synth
from oracle import orb_oracle as orc
frames = synth.frames(6, 480, 752, 2000, "dense"); frames[3] = synth.frame(480, 752, 9, "natural")
ex = ORBExtractor(1000, 1.2, 8, 20, 7, max_batch=4)
n, kps, desc = ex.extract_batch(frames)
oc = orc.Extractor(1000, 1.2, 8, 20, 7)
for b in range(6):
    ok, od = oc(frames[b])
    k1, d1 = ex(frames[b])
    bad_b = np.nonzero((desc[b,:n[b]] != od).any(1))[0]; bad_s = np.nonzero((d1 != od).any(1))[0]
    print(b, n[b], len(ok), 'batch-vs-oracle bad rows', len(bad_b), bad_b[:5], 'octaves', np.unique(ok['octave'][bad_b]) if len(bad_b) else '', '| single-vs-oracle bad', len(bad_s), bad_s[:5])
