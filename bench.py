#!/usr/bin/env python
"""bench.py -- throughput of the helmnet inference inner loop (BASELINE.json metric).

    python bench.py --gpus 1 --steps 30 --warmup 5                     # our arm, 1 B200
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --steps 5 --warmup 1              # the reference's CPU path (oracle port)

A "step" is ONE solver iteration (UNet update + spectral residual + residual norm) over the whole batch.
Workload (config.workload): 256x256 synthetic heterogeneous sos maps, batch 256 per GPU, point source at
[30,128], jcp_paper weights (the slim re-save of the shipped checkpoint in tests/golden).
Metric: Mpoint-iterations/s = batch * N^2 * steps / seconds, aggregated over all ranks.
  value : hn_run timed with CUDA events on the launch stream, all inputs resident in HBM.
  e2e   : IterativeSolver.forward() from pinned HOST sos maps to pinned HOST wavefield + rmse history.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(GOLD, "jcp_paper_trained_weights_slim.ckpt")
FLOP_PER_POINT = 16103.25          # 2 x 8051.625 MAC, dense count, per point-iteration
BYTES_PER_POINT_SPECTRAL = 36.0    # read wf 8 (updated in place by the UNet epilogue: d_wf never
#                                    exists in HBM) ... see DESIGN.md; broadcast source


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nme, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(n: int, batch: int, iters: int, warmup: int, threads: int):
    """The reference's CPU PyTorch path, restated (oracle/helmnet_oracle.py), timed on the host cores."""
    import torch

    from helmnet_b200.checkpoint import load_checkpoint
    from helmnet_b200.synthetic import synthetic_sos
    from oracle.helmnet_oracle import Oracle, point_source, rmse, zero_states

    torch.set_num_threads(threads)
    sd = load_checkpoint(CKPT)["state_dict"]
    w = {k[2:]: v for k, v in sd.items() if k.startswith("f.")}
    orc = Oracle(w, n)
    orc.set_source(point_source(n, [30, n // 2]))
    sos = synthetic_sos(batch, n, seed=1)
    k_sq, wf = orc.get_initials(sos)
    states = zero_states(batch, n)
    res = orc.residual(wf, k_sq)
    for _ in range(warmup):
        wf, res, states = orc.single_step(wf, k_sq, res, states)
    t0 = time.perf_counter()
    for _ in range(iters):
        wf, res, states = orc.single_step(wf, k_sq, res, states)
        _ = rmse(res)
    dt = time.perf_counter() - t0
    return batch * n * n * iters / dt / 1e6, dt / iters * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, b = args.n, args.cpu_batch
    val, ms = cpu_port_throughput(n, b, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "Mpoint-iterations/s", "value": val, "unit": "Mpoint-iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}x{n} synthetic sos, batch {args.batch} per GPU (reference arm: bounded sample, batch {b} per step)",
                   "n": n, "batch_per_gpu": args.batch},
        "cpu_baseline": {"value": val, "unit": "Mpoint-iterations/s", "cores": threads, "kind": "port",
                         "sample": f"{n}x{n}, batch {b}, {args.steps} iterations after {args.warmup} warm-up, torch CPU fp32"},
        "e2e": {"value": val, "unit": "Mpoint-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from helmnet_b200 import IterativeSolver
    from helmnet_b200.sharding import gather_results
    from helmnet_b200.synthetic import synthetic_sos

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; helmnet_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K, W = args.n, args.steps, args.warmup
    b_local = args.batch if args.scaling == "weak" else args.batch // world
    b_total = b_local * world

    solver = IterativeSolver.load_from_checkpoint(CKPT, strict=False, test_data_path=None)
    solver.freeze()
    solver.to(dev)
    solver.set_domain_size(n, source_location=[30, n // 2])
    # distinct maps per rank; 32 distinct maps tiled over the batch keep host generation short
    base = synthetic_sos(min(b_local, 32), n, seed=1 + rank)
    sos_host = base.repeat((b_local + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:b_local].contiguous().pin_memory()
    sos_dev = sos_host.to(dev, non_blocking=True)
    lib, ptr, stream = solver.lib, solver._ptr, solver._stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: `value` ------------------------------------------------
    ctx = solver._ensure_ctx(b_local)
    lib.check(lib.hn_reset(ctx, ptr(sos_dev), b_local, stream()), "hn_reset")
    rmse_buf = torch.empty(max(K, W), b_local, device=dev)
    lib.check(lib.hn_run(ctx, W, ptr(rmse_buf), ptr(None), ptr(None), ptr(None), stream()), "hn_run(warmup)")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    BAD = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")
    remeasured = False
    for attempt in range(2):
        launches0 = lib.hn_launch_count(ctx)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0.record()
        lib.check(lib.hn_run(ctx, K, ptr(rmse_buf), ptr(None), ptr(None), ptr(None), stream()), "hn_run")
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        launches = lib.hn_launch_count(ctx) - launches0
        # a run that saw a hardware / thermal slowdown is rejected and measured again, once (sw_power_cap is kept and noted)
        redo = torch.tensor([1 if (rank == 0 and attempt == 0 and any(r in BAD for r in clocks["reasons"])) else 0], device=dev)
        if world > 1:
            dist.all_reduce(redo, op=dist.ReduceOp.MAX)
        if int(redo.item()) == 0:
            break
        remeasured = True
        time.sleep(2.0)
    if clocks is not None:
        clocks["remeasured"] = remeasured
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = b_total * n * n * K / (ms_max * 1e-3) / 1e6

    # ---------------- per-stage and per-kernel device time (CUDA events on the launch stream) -------------
    import ctypes as C
    stage = (C.c_float * 2)()
    st_ms = [0.0, 0.0]
    reps = 5
    for _ in range(reps):
        lib.check(lib.hn_profile_iteration(ctx, stage, stream()), "hn_profile_iteration")
        st_ms[0] += stage[0] / reps
        st_ms[1] += stage[1] / reps
    # algorithmic HBM bytes per level-0 point and launch: activations are NHWC8 fp32 = 32 B
    fused = solver._engine >= 2 and n <= 256 and (n in (32, 64, 128, 256) or os.environ.get("HELMNET_TCF_ANY_WIDTH", "1") != "0")
    if fused:   # engine 2: one kernel per DoubleConv; the intermediate 8-channel tensor never reaches HBM
        kernel_table = [
            (9, "dconv_tcf_kernel<A8_B8,OUTC> (decode[0] 16->8->8 + outc, level 0; reads up 32 + skip 32, writes d_wf 8)", 32 + 32 + 8),
            (7, "dconv_tcf_kernel<A8_B2,STORE> (enc[0].conv_signal 10->8->8, level 0)", 32 + 8 + 32),
            (8, "dconv_tcf_kernel<A8_B2,STORE2> (enc[0].conv_state 10->2->2, level 0)", 32 + 8 + 8),
            (6, "dconv_tcf_kernel<INC,STORE> (inc 6->8->8, level 0; reads wf 8 + res 8)", 8 + 8 + 32),
            (2, "down_tcr_kernel (enc[0].down, level 0 -> 1)", 32 + 8),
            (3, "up_tcr_kernel (up[0], level 1 -> 0)", 8 + 32),
            (4, f"spectral_rows{n}_kernel" if n in (256, 512, 1024) else "spectral_rows_kernel", 8 + 8),
            (5, f"spectral_cols{n}_kernel" if n in (256, 512, 1024) else "spectral_cols_kernel", 8 + 8 + 4 + 8),
        ]
    else:
        kernel_table = [
            (0, "conv3x3_tcr_kernel<SRC_A8> (inc conv #2, 8->8, level 0)", 32 + 32),
            (1, "conv3x3_tcr_kernel<SRC_A8_B8> (decode[0] conv #1, 16->8, level 0)", 64 + 32),
            (2, "down_tcr_kernel (enc[0].down, level 0 -> 1)", 32 + 8),
            (3, "up_tcr_kernel (up[0], level 1 -> 0)", 8 + 32),
            (4, f"spectral_rows{n}_kernel" if n in (256, 512, 1024) else "spectral_rows_kernel", 8 + 8),
            (5, f"spectral_cols{n}_kernel" if n in (256, 512, 1024) else "spectral_cols_kernel", 8 + 8 + 4 + 8),
        ]
    kern_ms = {}
    one = C.c_float()
    for which, _, _ in kernel_table:
        lib.check(lib.hn_profile_layer(ctx, which, 10, C.byref(one), stream()), "hn_profile_layer")
        kern_ms[which] = float(one.value)

    # ---------------- end to end through the public API: `e2e` ----------------------------------------
    wf_host = torch.empty(b_local, 2, n, n).pin_memory()
    rm_host = torch.empty(K, b_local).pin_memory()

    def solve_e2e():
        out = solver.forward(sos_host.to(dev, non_blocking=True), num_iterations=K, return_residuals=False)
        wf, rm = out["wavefields"][0], out["residual_rmse"]
        if world > 1:   # NCCL only gathers results and residual histories
            gathered = gather_results(wf, rm, dst=0)
        wf_host.copy_(wf, non_blocking=True)
        rm_host.copy_(rm, non_blocking=True)

    with torch.no_grad():
        solve_e2e()   # warm-up (allocations, graph already built)
        barrier()
        e0.record()
        solve_e2e()
        e1.record()
        barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_val = b_total * n * n * K / (float(ms_e2e.item()) * 1e-3) / 1e6
    h2d = b_local * n * n * 4 / K
    d2h = (b_local * 2 * n * n * 4 + K * b_local * 4) / K

    # ---------------- BASELINE.json's second figure: wall time until every sample's residual RMSE < 1e-3 ----------
    ttr = None
    if args.residual_iters > 0:
        with torch.no_grad():
            hist = solver.forward(sos_dev, num_iterations=args.residual_iters, return_residuals=False)["residual_rmse"]
            below = (hist.max(dim=1).values < 1e-3).nonzero()
            k_star = int(below[0].item()) if below.numel() else -1
            kk = torch.tensor([k_star], device=dev)
            reached = torch.tensor([1 if k_star >= 0 else 0], device=dev)
            if world > 1:
                dist.all_reduce(kk, op=dist.ReduceOp.MAX)
                dist.all_reduce(reached, op=dist.ReduceOp.MIN)
            if int(reached.item()) == 1:
                k_all = int(kk.item())
                barrier()
                e0.record()
                out = solver.forward(sos_host.to(dev, non_blocking=True), num_iterations=k_all + 1, return_residuals=False)
                wf_host.copy_(out["wavefields"][0], non_blocking=True)
                e1.record()
                barrier()
                t_res = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
                ttr = {"reached": True, "iteration_index": k_all, "ms": float(t_res.item()),
                       "definition": "forward() from host sos until max over the batch of test_loss_function RMSE < 1e-3, incl. H2D/D2H"}
            else:
                ttr = {"reached": False, "iteration_index": None, "ms": None, "iterations_tried": args.residual_iters}
            # per-sample view (this rank): the reference network itself does not bring every out-of-distribution map below
            # 1e-3 within the cap (oracle run of the same maps: transient excursions of the residual, then slow decay)
            first = (hist < 1e-3).float().argmax(dim=0)
            ok_s = (hist < 1e-3).any(dim=0)
            ttr["samples_reached_frac"] = float(ok_s.float().mean().item())
            ttr["median_iteration_index_of_reached"] = int(first[ok_s].median().item()) if bool(ok_s.any()) else None
            ttr["ms_per_iteration"] = ms_max / K

    # ---------------- config[0] of BASELINE.json: the README lens (one 256x256 map), time until RMSE < 1e-3 ------------
    readme = None
    if rank == 0 and args.residual_iters > 0 and n == 256:
        import numpy as np
        lens = np.ones((n, n), np.float32)
        lens[100:170, 30:240] = np.tile(np.linspace(2, 1, 210), (70, 1))       # README.md:65-67
        lens_host = torch.from_numpy(lens)[None, None].contiguous().pin_memory()
        with torch.no_grad():
            h1 = solver.forward(lens_host.to(dev), num_iterations=120, return_residuals=False)["residual_rmse"][:, 0]
            bel = (h1 < 1e-3).nonzero()
            if bel.numel():
                k1 = int(bel[0].item())
                solver.forward(lens_host.to(dev), num_iterations=k1 + 1, return_residuals=False)     # warm (graph for B = 1)
                torch.cuda.synchronize()
                e0.record()
                o1 = solver.forward(lens_host.to(dev, non_blocking=True), num_iterations=k1 + 1, return_residuals=False)
                w1 = o1["wavefields"][0].cpu()
                e1.record()
                torch.cuda.synchronize()
                readme = {"iteration_index": k1, "ms": e0.elapsed_time(e1), "batch": 1,
                          "definition": "README.md:62-70 example: forward() from the host sos map until RMSE < 1e-3, wavefield back on the host"}
                if not args.no_cpu_baseline:
                    # the same solve on the reference's CPU path (oracle port), all host cores
                    from helmnet_b200.checkpoint import load_checkpoint
                    from oracle.helmnet_oracle import Oracle, point_source, zero_states
                    torch.set_num_threads(os.cpu_count() or 1)
                    sd = load_checkpoint(CKPT)["state_dict"]
                    orc = Oracle({k[2:]: v for k, v in sd.items() if k.startswith("f.")}, n)
                    orc.set_source(point_source(n, [30, n // 2]))
                    t0 = time.perf_counter()
                    k_sq, wf = orc.get_initials(torch.from_numpy(lens)[None, None])
                    st_, rs = zero_states(1, n), None
                    rs = orc.residual(wf, k_sq)
                    for _ in range(k1 + 1):
                        wf, rs, st_ = orc.single_step(wf, k_sq, rs, st_)
                    readme["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
                    readme["cpu_cores"] = os.cpu_count() or 1
        # the solver goes back to the benchmark batch lazily (next forward) -- nothing else runs after this point

    if rank == 0:
        peaks = load_peaks()
        pts = b_local * n * n
        t_unet, t_spec = st_ms[0] * 1e-3, st_ms[1] * 1e-3
        tf32_peak = peaks["bf16_tflops"] / 2.0
        ach_tf = FLOP_PER_POINT * pts / t_unet / 1e12 if t_unet > 0 else 0.0
        ach_gb = BYTES_PER_POINT_SPECTRAL * pts / t_spec / 1e9 if t_spec > 0 else 0.0
        # DRAM bytes per launch from the committed ncu capture of this very configuration (null for any other)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r1_dram_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj["config"] == {"n": n, "batch_per_gpu": b_local} and fused:
                traffic = {int(k): v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj["kernels"].items()}
        roof = []
        for which, name, bpp in kernel_table:
            ms_k = kern_ms[which]
            gbs = bpp * pts / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
            roof.append({"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                         "traffic": traffic.get(which), "algorithmic_bytes": bpp * pts, "kernel": name, "ms_per_launch": ms_k, "algorithmic_bytes_per_point": bpp,
                         "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json hbm_gbs, burst copy)"})
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, _ = cpu_port_throughput(n, args.cpu_batch, args.cpu_iters, 2, threads)
            cpu = {"value": v, "unit": "Mpoint-iterations/s", "cores": threads, "kind": "port",
                   "sample": f"{n}x{n}, batch {args.cpu_batch}, {args.cpu_iters} iterations after 2 warm-up, torch CPU fp32 "
                             "(oracle/helmnet_oracle.py)"}
        line = {
            "metric": "Mpoint-iterations/s", "value": value, "unit": "Mpoint-iterations/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{n}x{n} synthetic heterogeneous sos maps, batch {b_local} per GPU ({b_total} total), "
                                   f"point source [30,{n // 2}], jcp_paper weights", "n": n, "batch_per_gpu": b_local,
                       "global_batch": b_total, "parallelism": f"batch-sharded x{world}, no in-loop collective",
                       "l2": "working set per iteration (>4 GB) exceeds the 126 MB L2; no flush needed",
                       "e2e": "one forward() of `steps` iterations incl. H2D of sos and D2H of wavefield+rmse; bytes are per-step averages"},
            "e2e": {"value": e2e_val, "unit": "Mpoint-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "kernels_per_iteration": int(lib.hn_kernels_per_iteration(ctx)),
            "clocks": clocks,
            "roofline": roof[0],
            "roofline_kernels": roof[1:],
            "roofline_stage_unet": {"bound": "tensor", "achieved": ach_tf, "peak": tf32_peak, "unit": "TFLOP/s",
                                    "frac": ach_tf / tf32_peak if tf32_peak else None, "stage_ms": st_ms[0],
                                    "note": "16103.25 FLOP/point over the whole conv stack vs the dense TF32-rate peak "
                                            f"({peaks['source']} bf16 burst / 2); the C_out=8 GEMMs are HBM-bound, see `roofline`",
                                    "engine": solver._engine},
            "roofline_stage_spectral": {"bound": "hbm", "achieved": ach_gb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                        "frac": ach_gb / peaks["hbm_gbs"], "stage_ms": st_ms[1],
                                        "algorithmic_bytes_per_point": BYTES_PER_POINT_SPECTRAL},
            "cpu_baseline": cpu,
            "ms_to_residual_1e-3": ttr,
            "readme_lens_ms_to_residual_1e-3": readme,
            "final_rmse_max": float(rmse_buf[K - 1].max().item()),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--batch", type=int, default=256, help="batch per GPU (weak) or total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-iters", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--residual-iters", type=int, default=1000, help="iteration cap of the time-to-residual-1e-3 run (0: skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
