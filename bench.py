#!/usr/bin/env python
"""bench.py -- headline benchmark of the integrating-sphere hot path (BASELINE.json config C3).

A "step" = one pass of the hot path over one batch of synthetic source rays PER GPU:
trace (source -> multi-bounce loop with the CustomMirror BRDF) + 180x90 flux map + (N>1) one all-reduce.
Metric = ray-bounces/s summed over all GPUs ("scaling": "weak": each GPU gets --rays rays per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--map direction|line] [--impl reference]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray_bounces_per_s_fluxmap"
UNIT = "ray-bounces/s"
FLOP_PER_BOUNCE = 100.0          # Lambert + Gaussian-roughness configuration
REF_RECORDED = 3.7e6             # reference's own recorded rate, author's PC, <=4 threads


def workload_scene(mod):
    # C3: fluxAtObserverFast.C:33-41 scene + per-bounce spec/diffuse mixture of nonLambertianFlux.C:147-211
    return mod.scene(theta_max=170.0, world_half=300.0, reflectance=0.99, roughness=0.01, max_bounces=50000,
                     brdf_kind=1, brdf_param=(0.3, 0.4, 0.6, 0.0))


def workload_config(args, mode_name):
    return {"workload": "C3 nonLambertianFlux: CustomMirror BRDF (0.3,0.4,0.6), theta_max=170, rho=0.99, "
                        "sigma=0.01, src(-60,0,-75) dir(5,0,0), 180x90 map",
            "rays_per_gpu_per_step": args.rays, "map_mode": mode_name, "seed": 4357,
            "l2": "working set (32 B/ray record buffer, 8 GiB per 2^28-ray batch) exceeds the 126 MB L2; "
                  "the RNG is counter-based, there is no input to cache"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self, t0=None, t1=None):
        """Samples taken inside the timed region [t0, t1] (host clock); a region shorter than the sampling period falls back
        to the samples closest to it (the sampler runs from before the warm-up on)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        rows = [r for r in self.rows if len(r) > 3 and r[2].isdigit()]
        if t0 is not None and rows:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.25]
            rows = inside if inside else sorted(rows, key=lambda r: abs(r[0] - t1))[:2]
        rows = [r[1:] for r in rows]
        sm = sorted(int(r[1]) for r in rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  ROOT + ROBAST cannot be built here
    (DESIGN.md), so this times the oracle's double-precision restatement on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    O.build()
    mode = O.MAP_DIRECTION if args.map == "direction" else O.MAP_LINE
    sc, src, mp = workload_scene(O), O.source(), O.map_spec(mode=mode)
    cores = O.lib().orc_num_threads()
    sample = args.ref_rays
    for w in range(args.warmup):
        O.fluxmap(sc, src, max(sample // 10, 1000), mp, seed=4357, ray_id0=w * sample, prec=O.F64, n_threads=0)
    bounces = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        _, st = O.fluxmap(sc, src, sample, mp, seed=4357, ray_id0=(args.warmup + s) * sample, prec=O.F64, n_threads=0)
        bounces += st["n_bounces"]
    dt = time.perf_counter() - t0
    v = bounces / dt
    cfg = workload_config(args, args.map)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} rays per step of the same workload ({bounces} bounces in {dt:.1f} s); "
                                       "ROOT+ROBAST are not installable here, this is the FP64 oracle restatement"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rays_per_s": sample * args.steps / dt, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    # stdout carries exactly ONE line, the JSON: libraries that write to fd 1 on their own (NCCL prints its version banner
    # at the first communicator when NCCL_DEBUG is set on the box) are sent to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import altair_raytracing_b200 as A
    from altair_raytracing_b200.distributed import ShardedTracer

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = A.MAP_DIRECTION if args.map == "direction" else A.MAP_LINE
    sc, src, mp = workload_scene(A), A.source(), A.map_spec(mode=mode)
    ctx = A.Context([local])
    tr = ShardedTracer(ctx, sc, src, mp, seed=4357, device=local)
    R = args.rays
    nb = mp.n_theta * mp.n_phi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total = torch.zeros(8, dtype=torch.int64, device=tr.device)
    step_no = [0]

    def one_step():
        # weak scaling: the job traces world*R rays per step, this rank its shard of them
        buf = tr.step_device(world * R, ray_id0=step_no[0] * world * R)
        step_no[0] += 1
        return buf

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    l0 = ctx.launches
    t_region0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        buf = one_step()
        total += buf[nb:nb + 8]
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=tr.device)
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clk = clocks.stop(t_region0, time.time()) if rank == 0 else None
    launches = ctx.launches - l0 + (args.steps if world > 1 else 0)   # + the all-reduce kernels
    dt = ms.item() * 1e-3
    st = total.cpu().numpy()
    rays_done, bounces = int(st[0]), int(st[5])
    value = bounces / dt

    # ---- e2e: the public blocking call, host results every step (params H2D, map+stats D2H)
    barrier()
    t0 = time.perf_counter()
    e2e_bounces = 0
    for _ in range(args.steps):
        counts, stats = tr.step(world * R, ray_id0=step_no[0] * world * R)
        step_no[0] += 1
        e2e_bounces += int(stats[0, 5])
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=tr.device)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = e2e_bounces / e2e_dt.item()
    h2d = C.sizeof(A.Scene) + C.sizeof(A.Source) + C.sizeof(A.MapSpec) + 48
    d2h = (nb + 8) * 8

    # ---- per-kernel timing for the roofline (CUDA events inside the library around the trace launches)
    _, kst = ctx.trace_fluxmap(sc, src, R, mp, seed=4357, ray_id0=step_no[0] * world * R + rank * R)
    kst = kst[0]
    peak = ctx.measure_fp32_peak()
    n_batches = -(-R // (1 << 28))
    achieved = FLOP_PER_BOUNCE * kst["n_bounces"] / kst["t_trace_s"] * 1e-12
    roofline = {"bound": "fp32", "kernel": "k_trace<rough,CustomMirror>", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE 2^28-ray launch of this kernel, from the
                # `ncu --set full` capture in profiles/r01_ncu_final.md (algorithmic: 2^28 rays x 32 B = 8.590e9 B)
                "traffic": 9.4184e9, "traffic_unit": "B per 2^28-ray launch (ncu, profiles/r01_ncu_final.md)",
                "peak_source": "FFMA-chain probe measured live on this GPU (MEASURED_PEAKS.json has no FP32 number); "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "flop_per_bounce": FLOP_PER_BOUNCE, "launches": n_batches,
                "avg_launch_ms": kst["t_trace_s"] * 1e3 / n_batches, "map_ms_per_launch": kst["t_map_s"] * 1e3 / n_batches,
                "bounces_per_s_kernel": kst["n_bounces"] / kst["t_trace_s"],
                "issue_slots_busy_ncu": 0.823, "active_lanes_ncu": 27.6, "fma_pipe_ncu": 0.54, "alu_pipe_ncu": 0.52,
                "warp_instructions_per_bounce_ncu": 296,
                "hbm_bytes_per_bounce_algorithmic": 32.0 * kst["n_rays"] / kst["n_bounces"]}
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.map),
            "rays_per_s": rays_done / dt, "bounces_per_ray": bounces / max(rays_done, 1),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clk,
            "reference_recorded": {"value": REF_RECORDED, "unit": UNIT, "note": "BASELINE.md, author's PC, <=4 threads"}}
    if world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyoracle as O
            O.build()
            osc, osrc = workload_scene(O), O.source()
            omp_ = O.map_spec(mode=O.MAP_DIRECTION if args.map == "direction" else O.MAP_LINE)
            cores = O.lib().orc_num_threads()
            sample = args.ref_rays if args.map == "direction" else max(args.ref_rays // 20, 1000)   # LINE is brute force on the CPU
            t0 = time.perf_counter()
            _, ost = O.fluxmap(osc, osrc, sample, omp_, seed=4357, prec=O.F64, n_threads=0)
            cdt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ost["n_bounces"] / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} rays of the same workload, FP64 oracle restatement (ROOT+ROBAST cannot "
                                              f"run here), {cdt:.1f} s"}
            s1 = max(sample // (2 * max(cores, 1)), 1000)          # the same on ONE core, ~half the time again
            t0 = time.perf_counter()
            _, ost1 = O.fluxmap(osc, osrc, s1, omp_, seed=4357, prec=O.F64, n_threads=1)
            line["cpu_baseline"]["value_1core"] = ost1["n_bounces"] / (time.perf_counter() - t0)
        except Exception as e:          # the CPU leg must never cost the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=1_000_000_000, help="rays per GPU per step (C3: 1e9)")
    ap.add_argument("--map", choices=["direction", "line"], default="direction")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--ref-rays", type=int, default=20_000_000, help="CPU sample size (rays)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
