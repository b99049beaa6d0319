"""Test-set evaluation driver: the role of the reference's evaluate.py (`Evaluation.results_on_test_set`,
evaluate.py:27-29 -> Lightning `Trainer.test` -> `test_step` / `test_epoch_end`, hybridnet.py:299-330) without Lightning.

    python -m helmnet_b200.evaluate --model_checkpoint trained_models/jcp_paper_trained_weights.ckpt \\
        --test_set sos_maps.pt --gpu 0

`--test_set` is a torch file holding a float tensor [n, 1, N, N] (or [n, N, N]) of sound-speed maps; the reference's
pickled `EllipsesDataset` objects need its dataloaders module and are not read here.
Outputs (same names/shapes the reference writes and produce_figures.py reads):
    results/evolution_of_model_RMSE_on_test_set.npy   [n, max_iterations]
    results/evolution_of_wavefields_on_test_set.npy   [n, max_iterations, 2, N, N]
"""
from __future__ import annotations

import argparse

import torch

from .solver import IterativeSolver


def get_model(path: str, domain_size=None, source_location=None) -> IterativeSolver:
    """evaluate.py:48-71: rebuild the solver from the checkpoint's hparams and copy the `f` weights."""
    model = IterativeSolver.load_from_checkpoint(path, strict=False, test_data_path=None)
    hp = dict(model.hparams)
    if domain_size is not None:
        hp["domain_size"] = domain_size
    if source_location is not None:
        hp["source_location"] = source_location
    new_model = IterativeSolver(**hp)
    new_model.f.load_state_dict(model.f.state_dict())
    new_model.set_laplacian()
    new_model.set_source()
    new_model.freeze()
    return new_model


def results_on_test_set(model: IterativeSolver, sos_maps: torch.Tensor, batch_size: int = 32, out_dir: str = "results",
                        max_iterations=None):
    if sos_maps.dim() == 3:
        sos_maps = sos_maps.unsqueeze(1)
    if max_iterations is not None:
        model.hparams.max_iterations = int(max_iterations)
    outputs = []
    with torch.no_grad():
        for i in range(0, sos_maps.shape[0], batch_size):
            out = model.test_step(sos_maps[i: i + batch_size].to(model.device).float(), i // batch_size)
            outputs.append({"losses": out["losses"].cpu(), "wavefields": [w.cpu() for w in out["wavefields"]]})
    return model.test_epoch_end(outputs, out_dir=out_dir)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_checkpoint", type=str, default="checkpoints/trained_weights.ckpt")
    ap.add_argument("--test_set", type=str, required=True)
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--max_iterations", type=int, default=None)
    args = ap.parse_args()
    model = get_model(args.model_checkpoint)
    model.to(f"cuda:{args.gpu}")
    sos = torch.load(args.test_set)
    losses = results_on_test_set(model, sos, args.batch_size, max_iterations=args.max_iterations)
    print("final residual RMSE: mean %.3e  max %.3e" % (losses[:, -1].mean(), losses[:, -1].max()))


if __name__ == "__main__":
    main()
