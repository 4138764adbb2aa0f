import sys, os; sys.path.insert(0,'/root/repo')
import torch
from chap_b200 import ops, _lib
dev="cuda"
def run(cin,cout,s,n,label):
    x = ops.cl(torch.randn(n,cin,s,s,device=dev))
    w = (torch.randn(cout,cin,3,3,device=dev)/ (cin*9)**0.5).requires_grad_(True)
    y = ops.conv(x,w,None,_lib.CONV_K3)
    gy = torch.randn_like(y)
    for _ in range(3): torch.autograd.grad(y,w,gy,retain_graph=True)
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    for _ in range(10): torch.autograd.grad(y,w,gy,retain_graph=True)
    fam=_lib.timing_report(); _lib.timing_enable(False)
    f=fam.get("conv_tc_wgrad") or fam.get("conv_thin_wgrad") or fam.get("conv_simt_wgrad")
    print(label, cin,cout,s,n, "us/launch %.1f"%(f["ms"]/f["launches"]*1e3), "GB/s(alg) %.0f"%(f["bytes"]/f["ms"]/1e6), "TF/s %.1f"%(f["flops"]/f["ms"]/1e9), flush=True)
tag=os.environ.get("CHAP_THIN_MAX","0")+("" if not os.environ.get("CHAP_NO_ROW_REUSE") else "-noreuse")
for (ci,co,s,n) in ((16,16,256,12),(32,16,256,12),(32,32,128,12),(64,32,128,12),(64,64,64,12)):
    run(ci,co,s,n,"dbg="+tag)
