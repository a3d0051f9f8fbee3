"""developer aid: ONE sharded-plan product of 2^LOG2 x 2^LOG2 limbs on one GPU (for ncu launch lists)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, mpir_fft_b200 as M
from mpir_fft_b200.sharded import ShardedMul
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
torch.cuda.set_device(0); M.init(0)
n = 1 << log2
d, w = bench.SHARDED_PARAMS[log2]
a = bench.splitmix64_dev(torch, 1, n, torch.device("cuda", 0)); b = bench.splitmix64_dev(torch, 2, n, torch.device("cuda", 0))
sm = ShardedMul(n, n, d, w, cuda=True, single=True)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    sm.multiply(a.data_ptr(), b.data_ptr())
torch.cuda.synchronize()
print("done")
