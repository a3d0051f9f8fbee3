"""developer aid: per-CTA phase timeline of the fused tile kernel for each pass of a cfg2 product"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpir_fft_b200 as M
M.init(0)
L = M.lib()
L.mfft_dev_tile_timing.argtypes = [C.c_void_p]
n = 1 << 20
plan = M.MulPlan(n, n, 14, 1)
a = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda")
b = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda")
r = torch.zeros(2 * n, dtype=torch.int64, device="cuda")
for _ in range(3):
    plan.exec_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), None)
torch.cuda.synchronize()
buf = torch.zeros(8 * 16384, dtype=torch.int64, device="cuda")
for ph, name, grids in ((0, "fwd", (2048, 1152, 1040, 1040)), (3, "inv", (1040, 1040, 1152, 2048, 1280))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    buf.zero_()
    L.mfft_dev_tile_timing(C.c_void_p(buf.data_ptr()))
    plan.exec_phase(ph, r.data_ptr(), a.data_ptr(), b.data_ptr(), None)
    torch.cuda.synchronize()
    L.mfft_dev_tile_timing(None)
    e0.record()
    for _ in range(10):
        plan.exec_phase(ph, r.data_ptr(), a.data_ptr(), b.data_ptr(), None)
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.1f us per transform (10 back-to-back, un-instrumented)" % (name, e0.elapsed_time(e1) * 100))
    allt = buf.cpu().numpy().reshape(-1, 8)
    off = 0
    for gi, g in enumerate(grids):
        t = allt[off:off + g]; off += g
        assert (t[:, 0] > 0).all()
        t0 = t[:, 0].min()
        d = t[:, 1:5] - t[:, 0:4]
        print("%s launch %d: CTAs %d span %.1f us | mean us: issue %.2f wait %.2f ops %.2f store %.2f residency %.2f | p90: %s" % (
            name, gi, g, (t[:, 4].max() - t0) / 1e3, d[:, 0].mean() / 1e3, d[:, 1].mean() / 1e3, d[:, 2].mean() / 1e3,
            d[:, 3].mean() / 1e3, (t[:, 4] - t[:, 0]).mean() / 1e3, np.round(np.percentile(d, 90, axis=0) / 1e3, 2)))
        starts = np.sort(t[:, 0] - t0) / 1e3
        print("      CTA start-time quantiles us:", np.round(np.percentile(starts, [0, 25, 50, 75, 100]), 1))
