import sys, numpy as np, ctypes as C, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common
from sugarscape_utilicalc_b200 import config_from_settings
from sugarscape_utilicalc_b200.engine import Engine, _ptr
n=int(sys.argv[1]); pol=bool(int(sys.argv[2])); win=int(sys.argv[3]) if len(sys.argv)>3 else 120
s, init = common.synthetic_case(n, pol)
cfg = config_from_settings(s, rng_mode="keyed")
e = Engine(cfg); e.set_option("window", win); common.init_world(e, init, "keyed")
e.step(5)
e.set_option("debug",1); e.set_option("reset_counters",0)
K=5
st=e.step(K)
out=np.zeros(38,dtype=np.int64); e._lib.ss_get_counters(e._h,_ptr(out),38)
d=out[6:]
print("n",n,"pol",pol,"window",out[5],"ms/step",e.kernel_times()["last_call"]/K,"passes/step",out[1]/K)
W=max(d[0],1); clk=1.9e9
print("windows-with-work",d[0],"agents listed/window",d[13]/W,"turns",d[9],"blocked",d[12])
print("per window: load %.1f us, +list %.1f us, total %.1f us; per warp: turn %.1f us wait %.1f us; per turn %.2f us"%(d[11]/W/clk*1e6, d[3]/W/clk*1e6, d[6]/W/clk*1e6, d[15]/W/8/clk*1e6, d[16]/W/8/clk*1e6, d[15]/max(d[9],1)/clk*1e6))
e.set_option("profile",1); e.set_option("reset_times",0); e.step(K); print({k:v/K for k,v in e.kernel_times().items()})
