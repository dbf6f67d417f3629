import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import test_gpu_conv_tc as t
rng=np.random.default_rng(0)
for (n,H,W,Cin,Cout,k) in [(64,25,20,64,128,3),(64,25,20,64,128,1),(64,25,20,64,32,3),(64,25,20,64,64,5),(64,25,20,16,32,3),(64,13,10,128,256,3)]:
    x=rng.standard_normal((n,H,W,Cin)).astype(np.float32); w=(rng.standard_normal((k,k,Cin,Cout))*0.05).astype(np.float32); b=np.zeros(Cout,np.float32)
    t.run_conv(0,3,x,w,b,n,H,W,Cin,Cout,k,1,0)
    print("shape",n,H,W,Cin,Cout,k,"entries",Cin//16*k*k, flush=True)
