import torch, sys
sys.path.insert(0,'/root/repo')
from handyrec_b200 import kernels as k
import oracle
dev=torch.device('cuda:0')
for (B,F,D) in [(5,3,8),(100,8,8),(257,3,32),(4096,26,16),(1000,7,32),(64,2,4),(1,1,4)]:
    x=torch.randn(B,F,D); w=torch.randn(D,1)*0.1; w0=torch.tensor([0.3])
    got=k.fm_fwd(x.to(dev),w.to(dev),w0.to(dev))
    torch.cuda.synchronize()
    err=(got.cpu()-oracle.fm(x,w,w0)).abs().max().item()
    print(B,F,D,'err',err,flush=True)
