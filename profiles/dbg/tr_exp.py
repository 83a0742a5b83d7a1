import time, ctypes as C, torch, sys, os
sys.path.insert(0,'/root/repo')
from kwage_b200 import capi
nf, cbits = 2048, 1 << 22
h_in = torch.empty(nf * cbits // 8, dtype=torch.uint8).pin_memory(); h_in.random_(0,256)
h_out = torch.empty(cbits * nf // 8, dtype=torch.uint8).pin_memory()
ptrs = (C.c_void_p * nf)(*[h_in.data_ptr() + j * (cbits // 8) for j in range(nf)])
for it in range(3):
    t0=time.time(); capi.check(capi.lib().kwg_transpose(0, ptrs, nf, cbits, C.c_void_p(h_out.data_ptr()))); print(os.environ.get('KWG_TR_BUDGET_MIB'), os.environ.get('KWG_TR_ONE_LANE'), round((time.time()-t0)*1e3,1),'ms')
# raw PCIe check
d=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t0=time.time(); d.copy_(h_in[:1<<30],non_blocking=True); torch.cuda.synchronize(); print('H2D 1GiB', round((time.time()-t0)*1e3,1))
t0=time.time(); h_out[:1<<30].copy_(d,non_blocking=True); torch.cuda.synchronize(); print('D2H 1GiB', round((time.time()-t0)*1e3,1))
