#!/usr/bin/env python3
"""All-pairs Hamming: GMatch/s of orbfe_hamming_allpairs_device at 40 000 x 40 000 (and other shapes given as NQxNT arguments), with a
bit-exact check against the CPU port on a sub-problem.  ORBFE_ALLPAIRS_POPC=1 selects the popc kernel instead of the int8 tensor-core one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoorbslam3_b200 import ORBExtractor, ORBMatcher
from oracle import orb_oracle as orc

ex = ORBExtractor(1000, 1.2, 8, 20, 7)
m = ORBMatcher(0.6, False, handle=ex._h)
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu"); g.manual_seed(7)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(40000, 40000)]
orc.build()
for nq, nt in shapes:
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, generator=g)
    t[torch.randint(0, nt, (nt // 10,), generator=g)] = q[torch.randint(0, nq, (nt // 10,), generator=g)]       # exact duplicates: distance 0 ties
    dq, dt = q.to(dev), t.to(dev)
    bi = torch.zeros(nq, dtype=torch.int32, device=dev); bd = torch.zeros_like(bi); sd = torch.zeros_like(bi)
    for _ in range(2): m.hamming_allpairs_device(dq, nq, dt, nt, bi, bd, sd, sync=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    st = torch.cuda.Stream()                               # a real stream handle (the default stream's handle is 0 = "the handle's own")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(K): m.hamming_allpairs_device(dq, nq, dt, nt, bi, bd, sd, stream=st.cuda_stream, sync=False)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    sub = min(nq, 1500)
    obi, obd, osd = orc.hamming_allpairs(q[:sub].numpy(), t.numpy())
    ok = np.array_equal(bi[:sub].cpu().numpy(), obi) and np.array_equal(bd[:sub].cpu().numpy(), obd) and np.array_equal(sd[:sub].cpu().numpy(), osd)
    tail = min(nq, 700)
    obi2, obd2, osd2 = orc.hamming_allpairs(q[nq - tail:].numpy(), t.numpy())
    ok2 = np.array_equal(bi[nq - tail:].cpu().numpy(), obi2) and np.array_equal(bd[nq - tail:].cpu().numpy(), obd2) and np.array_equal(sd[nq - tail:].cpu().numpy(), osd2)
    print("%d x %d: %.3f ms  %.1f GMatch/s  first %d rows exact: %s, last %d rows exact: %s" % (nq, nt, ms, nq * nt / ms / 1e6, sub, ok, tail, ok2))
