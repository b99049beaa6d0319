"""One training step (n_steps under autograd + loss.backward) at the reference's training configuration, for ncu launch lists
and timing: python tools/train_probe.py [n] [batch] [steps] [reps]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CKPT  # noqa: E402
from helmnet_b200 import IterativeSolver  # noqa: E402
from test_emu_train import unroll_case  # noqa: E402


def main():
    n, batch, steps, reps = (int(v) for v in (sys.argv[1:5] + ["96", "32", "10", "3"][len(sys.argv) - 1:]))
    s = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    s.train()
    s.to("cuda:0")
    s.set_domain_size(n, source_location=[n // 3, n // 2])
    wf, res, k_sq, states, _, _ = [[t.cuda() for t in c] if isinstance(c, list) else c.cuda() for c in unroll_case(n, batch, seed=7)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(reps):
        for p in s.f.parameters():
            p.grad = None
        s.f.set_states([h.clone() for h in states])
        e[0].record()
        out = s.n_steps(wf, k_sq, res, steps, True, True)
        loss = 1e4 * torch.cat(out["residuals"]).pow(2).mean()
        e[1].record()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        print(f"rep {rep}: n={n} batch={batch} steps={steps}: forward {e[0].elapsed_time(e[1]):.2f} ms, backward {e[1].elapsed_time(e[2]):.2f} ms, "
              f"loss {float(loss):.6e}", flush=True)


if __name__ == "__main__":
    main()
