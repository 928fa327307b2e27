import os, sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import gq_b200
from gq_b200 import _lib
from util import codebook
dev=torch.device('cuda',0)
cbt=torch.from_numpy(codebook(16,256)).to(dev)
n_chunks=1468652
xs=[torch.randn(n_chunks*16,device=dev)*0.01 for _ in range(4)]
codes=torch.empty(n_chunks,dtype=torch.uint8,device=dev); u=torch.empty(n_chunks,device=dev)
seg=torch.tensor([0,n_chunks],dtype=torch.int64,device=dev); ws=torch.empty(1<<20,dtype=torch.uint8,device=dev)
def run(i): _lib.call("gq_hsq_search", xs[i%4].data_ptr(), n_chunks, 16, cbt.data_ptr(), 256, codes.data_ptr(), 1, u.data_ptr(), seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), _lib.ALGO_TC, _lib.stream())
for i in range(5): run(i)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40): run(i)
e1.record(); torch.cuda.synchronize()
print("groups=%s flags=%s: %.4f ms"%(os.environ.get('GQ_TC_GROUPS','2'), os.environ.get('GQ_TC_FLAGS','0'), e0.elapsed_time(e1)/40))
