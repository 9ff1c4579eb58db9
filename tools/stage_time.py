"""Per-window stage times (cross-covariance, variance product) of the scoring pipeline at config C3 with the stages run back to
back (gpso_set_profile 1); used for same-box A/B runs of kernel variants.  usage: python tools/stage_time.py [label]"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from pygpso_b200 import backend
N,d,M=4096,10,1_200_000
X,y=bench.synthetic_training(N,d); theta=bench.fixed_theta(d)
s=backend.CudaBackend(device=0).open_session("Matern52",1,True); s.set_data(X,y); s.factorize(theta)
xc=torch.from_numpy(np.random.default_rng([bench.SEED,0]).random((M,d))).cuda()
st=torch.cuda.current_stream().cuda_stream
for _ in range(2): r=s.ucb_argmax_dev(xc.data_ptr(),M,bench.VARSIGMA,st)
s.set_profile(1)
acc=np.zeros(4)
for _ in range(3):
    r=s.ucb_argmax_dev(xc.data_ptr(),M,bench.VARSIGMA,st); acc+=s.last_timing_ms()
nw=s.last_windows()*3
print(sys.argv[1] if len(sys.argv)>1 else "", "per window ms: crosscov %.3f product %.3f ; argmax %d ucb %.15g" % (acc[1]/nw, acc[2]/nw, r[0], r[3]))
