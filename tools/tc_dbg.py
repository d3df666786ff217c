import sys, os
sys.path.insert(0, "gcn-song-embeddings_b200")
import torch, ps_native as nat
nat._ensure_device()
for (M,N,K) in [(128,128,32),(257,128,64),(1000,512,256)]:
    A=torch.randn(M,K,device="cuda"); B=torch.randn(N,K,device="cuda"); C=torch.empty(M,N,device="cuda")
    nat.gemm(A,B,C,M,N,K); torch.cuda.synchronize()
    err=float((C.double()-A.double()@B.double().t()).norm()/(A.double()@B.double().t()).norm())
    print(M,N,K,"rel err",err, flush=True)
