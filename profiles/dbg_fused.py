import sys; sys.path.insert(0,'.')
import torch, __graft_entry__ as e
pkg=e.load_package()
dev='cuda:0'; B,D,Tx,Ty=32,80,200,1000
x_m=torch.randn(B,D,Tx,device=dev); x_logs=0.3*torch.randn(B,D,Tx,device=dev)-0.5; z=torch.randn(B,D,Ty,device=dev)
xl=torch.full((B,),Tx,dtype=torch.int32,device=dev); yl=torch.full((B,),Ty,dtype=torch.int32,device=dev)
for i in range(3):
    pkg.fused_maximum_path(x_m,x_logs,z,xl,yl); torch.cuda.synchronize(); print('call',i,'ok')
