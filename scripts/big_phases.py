"""developer aid: per-phase device times of one sharded product on this rank (world 1)"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import big_sharded as B
import mpir_fft_b200 as M
from mpir_fft_b200.sharded import ShardedMul
log2, depth, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.cuda.set_device(0); M.init(0)
dev = torch.device("cuda", 0)
n = 1 << log2
a = B.splitmix64_dev(1, n, dev); b = B.splitmix64_dev(2, n, dev)
sm = ShardedMul(n, n, depth, w, cuda=True)
sm.multiply(a.data_ptr(), b.data_ptr()); torch.cuda.synchronize()
orig = sm._phase
times = {}
def timed(ph, which=0, d_in=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(ph, which, d_in); e1.record(); torch.cuda.synchronize()
    times.setdefault((ph, which), []).append(e0.elapsed_time(e1))
sm._phase = timed
sm.multiply(a.data_ptr(), b.data_ptr())
names = {0: "split+col FFT", 1: "unpack+row FFT", 2: "pointwise", 3: "row IFFT", 4: "col IFFT", 5: "unpack", 6: "recombine"}
for (ph, which), v in sorted(times.items()):
    print("phase %d (%s) operand %d: %.2f ms" % (ph, names[ph], which, sum(v)))
print("blocks of %d limbs; live blocks per slab %d" % (sm.lay.block_limbs, sm.lay.trunc_rows * sm.lay.n1cols))
