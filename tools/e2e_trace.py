import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from monoorbslam3_b200 import ORBExtractor, synth
from monoorbslam3_b200.extractor import KP_DTYPE
H,W,NF,B=480,752,1000,512
base=synth.frames(16,H,W,1000,'dense'); host=torch.from_numpy(np.concatenate([base]*(B//16))).pin_memory()
cap=NF+64
h_n=torch.zeros(B,dtype=torch.int32).pin_memory(); h_kps=torch.zeros((B,cap,7),dtype=torch.float32).pin_memory(); h_desc=torch.zeros((B,cap,32),dtype=torch.uint8).pin_memory()
out=(h_n.numpy(),h_kps.numpy().view(KP_DTYPE).reshape(B,cap),h_desc.numpy()); fr=host.numpy()
ex=ORBExtractor(NF,1.2,8,20,7,max_batch=B)
for _ in range(4): 
    sys.stderr.write('--- step\n'); ex.extract_batch(fr,cap=cap,out=out)
