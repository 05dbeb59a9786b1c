import json, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r)
from far_b200 import ops
from far_b200._lib import ENGINE_TCGEN05, ACT_NONE
def t(fn, it=20):
    for _ in range(3): fn()
    ts=[]
    for _ in range(it):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))*1e3
for (M,N,K) in ((303104,128,32),(303104,128,256),(151552,256,256)):
    x=torch.randn(M,K,device="cuda"); w=torch.randn(N,K,device="cuda")
    y=torch.empty(M,N,device="cuda")
    print(M,N,K, round(t(lambda: ops.linear(x,w,None,ACT_NONE,engine=ENGINE_TCGEN05,out=y)),1), "us")
# plain streaming references
a=torch.randn(303104,128,device="cuda"); b=torch.empty_like(a)
print("copy 155MB", round(t(lambda: b.copy_(a)),1), "us")
''' % ROOT
for dbg in ("0", "1", "2", "4", "6", "7"):
    for presplit in ("", "1"):
        env = dict(os.environ, FAR_TC_DBG=dbg)
        if presplit: env["FAR_TC_PRESPLIT"] = "1"
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"--- FAR_TC_DBG={dbg} presplit={bool(presplit)}"); print(out.stdout.strip()); print(out.stderr.strip()[-300:])
