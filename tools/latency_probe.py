import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from helmnet_b200 import IterativeSolver
s = IterativeSolver.load_from_checkpoint('/root/repo/tests/golden/jcp_paper_trained_weights_slim.ckpt', strict=False, test_data_path=None)
s.freeze(); s.to('cuda:0'); s.set_domain_size(256, source_location=[30, 128])
lens = np.ones((256, 256), np.float32); lens[100:170, 30:240] = np.tile(np.linspace(2, 1, 210), (70, 1))
x = torch.from_numpy(lens)[None, None].cuda()
for eng in (2, 1, 0):
    s.set_engine(eng)
    for B in (1, 4, 16):
        xb = x.repeat(B, 1, 1, 1)
        with torch.no_grad():
            s.forward(xb, num_iterations=60, return_residuals=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            o = s.forward(xb, num_iterations=200, return_residuals=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"engine {eng} B={B}: {dt/200*1e3:.3f} ms/iteration, rmse[52]={float(o['residual_rmse'][52,0]):.3e}")
