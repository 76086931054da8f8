#!/usr/bin/env python
"""bench.py -- ALS iterations/sec on the BASELINE.json workload (configs[1]: rank 64, 1M users x 100k
items, 100M synthetic implicit ratings), B200-native path vs the CPU restatement of the reference.

One "step" = one ALS iteration (item half-step + user half-step) over the whole rating set.
  value : iterations/sec with ratings + CSR already resident in HBM (CUDA events on the library stream)
  e2e   : iterations/sec through the one-shot C-ABI call pio_als_train with HOST buffers:
          H2D of the COO triplets, ingest (dedup + 2 CSR builds), K iterations, D2H of the factors.
See DESIGN.md for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (n_users, n_items, nnz, rank, implicit)
    "c2": (1_000_000, 100_000, 100_000_000, 64, True),
    "c1": (10_000, 1_000, 100_000, 10, False),
    "small": (100_000, 20_000, 5_000_000, 64, True),
}
LAMBDA, ALPHA, SEED = 0.01, 1.0, 3


def algorithmic_work(nu, ni, nnz, k, implicit):
    """BASELINE.md section 4 / SURVEY 8(d): bytes and flops per ALS iteration."""
    b = 16 * nnz + 8 * (nu + ni) * k + 8 * (nu + ni + 2)
    f_solve = 2 * nnz * (k * (k + 1) + 2 * k) + (nu + ni) * (k ** 3 / 3 + 2 * k * k)
    f_gram = (nu + ni) * k * (k + 1) if implicit else 0
    return b, f_solve, f_gram


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_from_profiles(workload):
    """dram__bytes_read+write of ONE launch of the dominant kernel, from the committed ncu --set full capture (null if none)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text())[workload]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def cpu_baseline_prepare(nu, ni, nnz, k, implicit, d_coo=None):
    """Load the oracle, generate (or take) the workload's ratings and build both CSR orientations - done once."""
    from oracle import als_oracle as o
    from pio_b200 import synth
    try:  # use a host-native build for the timed baseline when gcc is on the box
        import shutil
        if shutil.which("gcc"):
            so = o._SO.parent / "libals_oracle_native.so"
            subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
                            "-o", str(so), str(o._SRC), "-lm"], check=True)
            import ctypes
            o._lib = ctypes.CDLL(str(so))
            o._lib.oracle_num_threads.restype = ctypes.c_int
    except Exception:
        pass
    if d_coo is not None:
        u, i, r = d_coo
    else:
        u, i, r = synth.synth_ratings(nu, ni, nnz, seed=SEED, implicit=implicit)
    # duplicates are kept as separate ratings for the CPU sample (same arithmetic per rating)
    uptr, ucol, uval = o.csr_build(nu, u, i, r)
    iptr, icol, ival = o.csr_build(ni, i, u, r)
    uf = synth.synth_init_factors(min(nu, 1 << 16), k, SEED, 0)
    uf = np.ascontiguousarray(np.resize(uf, (nu, k)))
    itf = synth.synth_init_factors(min(ni, 1 << 16), k, SEED, 1)
    itf = np.ascontiguousarray(np.resize(itf, (ni, k)))
    prep = {"o": o, "nu": nu, "ni": ni, "implicit": implicit, "user": (uptr, ucol, uval), "item": (iptr, icol, ival),
            "uf": uf, "itf": itf, "yty": o.gram(uf) if implicit else None, "cores": o.num_threads()}

    def timed(side, stride):
        ptr, col, val = prep[side]
        src, dst, n = (itf, uf.copy(), nu) if side == "user" else (uf, itf.copy(), ni)
        t0 = time.perf_counter()
        o.half_step(ptr, col, val, src, dst, LAMBDA, implicit, ALPHA, prep["yty"], 0, n, stride)
        return time.perf_counter() - t0

    prep["timed"] = timed
    # pilot with a large stride: seconds per destination row of each side
    su, si = max(1, nu // 2000), max(1, ni // 500)
    prep["pilot"] = (su, timed("user", su), si, timed("item", si))
    return prep


def cpu_baseline_run(prep, target_s=12.0):
    """Time the oracle (CPU restatement of the MLlib algorithm) on a strided sample of destination rows of the SAME
    workload, sized to about target_s seconds, and scale to one full iteration."""
    o, timed = prep["o"], prep["timed"]
    su, tu, si, ti = prep["pilot"]
    su2 = max(1, int(su * tu / (target_s / 2)))
    si2 = max(1, int(si * ti / (target_s / 2)))
    tu = timed("user", su2)
    ti = timed("item", si2)
    t_gram = 0.0
    if prep["implicit"]:
        t0 = time.perf_counter()
        o.gram(prep["uf"])
        o.gram(prep["itf"])
        t_gram = time.perf_counter() - t0
    iter_s = tu * su2 + ti * si2 + t_gram
    return {"value": 1.0 / iter_s, "unit": "iterations/s", "cores": prep["cores"], "kind": "port",
            "sample": f"every {su2}th user row + every {si2}th item row of the full workload "
                      f"({tu + ti:.1f}s timed), scaled to one iteration; C/OpenMP restatement of MLlib ALS (not Spark)"}


def cpu_baseline_sample(nu, ni, nnz, k, implicit, target_s=12.0, d_coo=None):
    return cpu_baseline_run(cpu_baseline_prepare(nu, ni, nnz, k, implicit, d_coo=d_coo), target_s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PIO_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    nu, ni, nnz, k, implicit = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl_name = (f"ALS rank={k}, {nu} users x {ni} items, {nnz} synthetic "
               f"{'implicit' if implicit else 'explicit'} ratings")
    config = {"workload": wl_name, "lambda": LAMBDA, "alpha": ALPHA, "seed": SEED,
              "dedup": "sum" if implicit else "none",
              "l2": "inputs (1.6 GB of CSR + 282 MB of factors per iteration at c2) exceed the 126 MB L2",
              "parallelism": f"row-sharded x{args.gpus}, factor all-gather per half-iteration" if args.gpus > 1 else "1 GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        import pio_b200  # noqa: F401
        # one step = one bounded sample of the workload (same ratings, strided destination rows); the ratings and
        # both CSR orientations are built once; the per-step budget shrinks with the step count so that the whole
        # run stays within a few minutes
        prep = cpu_baseline_prepare(nu, ni, nnz, k, implicit)
        per_step = min(16.0, max(2.0, 90.0 / (args.warmup + args.steps)))
        vals = []
        cb = None
        for s in range(args.warmup + args.steps):
            cb = cpu_baseline_run(prep, target_s=per_step)
            if s >= args.warmup:
                vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "ALS iterations/sec", "value": v, "unit": "iterations/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64 accumulate / f32 storage", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import pio_b200  # noqa: F401
    from pio_b200 import native

    torch.cuda.set_device(local_rank)
    dev = local_rank
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(native.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    # ---- inputs resident in HBM ----------------------------------------------------------
    du = torch.empty(nnz, dtype=torch.int32, device="cuda")
    di = torch.empty(nnz, dtype=torch.int32, device="cuda")
    dr = torch.empty(nnz, dtype=torch.float32, device="cuda")
    native.synth_ratings_device(dev, nu, ni, nnz, SEED, implicit, 0, du.data_ptr(), di.data_ptr(), dr.data_ptr())
    dedup = native.DEDUP_SUM if implicit else native.DEDUP_NONE
    m = native.NativeALS(k, nu, ni, lam=LAMBDA, implicit=implicit, alpha=ALPHA, seed=SEED, device=dev,
                         world_size=world, world_rank=rank, nccl_id=nccl_id, init_mode=native.INIT_HASH)
    m.set_ratings_device(du.data_ptr(), di.data_ptr(), dr.data_ptr(), nnz, dedup=dedup)
    ingest_ms = m.stats()["last_ingest_ms"]
    nnz_eff = m.stats()["nnz"]
    m.run(max(args.warmup, 0))
    l0 = m.stats()
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    m.run(args.steps)
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()
    st = m.stats()
    dev_ms = st["last_run_ms"]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, st["last_solve_ms"], st["last_gram_ms"], st["last_comm_ms"], wall_s * 1e3],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, solve_ms, gram_ms, comm_ms, wall_ms = t.tolist()
    else:
        solve_ms, gram_ms, comm_ms, wall_ms = st["last_solve_ms"], st["last_gram_ms"], st["last_comm_ms"], wall_s * 1e3
    launches = st["kernel_launches"] - l0["kernel_launches"]
    solve_launches = st["solve_launches"] - l0["solve_launches"]
    value = args.steps / (dev_ms / 1e3)

    # ---- end to end through the C ABI with host buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        hu = torch.empty(nnz, dtype=torch.int32).pin_memory()
        hi = torch.empty(nnz, dtype=torch.int32).pin_memory()
        hr = torch.empty(nnz, dtype=torch.float32).pin_memory()
        hu.copy_(du)
        hi.copy_(di)
        hr.copy_(dr)
        out_u = torch.empty((nu, k), dtype=torch.float32).pin_memory()
        out_i = torch.empty((ni, k), dtype=torch.float32).pin_memory()
        m2 = native.NativeALS(k, nu, ni, lam=LAMBDA, implicit=implicit, alpha=ALPHA, seed=SEED, device=dev,
                              world_size=world, world_rank=rank, nccl_id=None if world == 1 else nccl_id_2(native, rank, world),
                              init_mode=native.INIT_HASH)
        m2.train(hu.numpy(), hi.numpy(), hr.numpy(), 1, dedup=dedup, out_user=out_u.numpy(), out_item=out_i.numpy())  # warm-up
        barrier()
        t0 = time.perf_counter()
        m2.train(hu.numpy(), hi.numpy(), hr.numpy(), args.steps, dedup=dedup, out_user=out_u.numpy(), out_item=out_i.numpy())
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        st2 = m2.stats()
        e2e = {"value": args.steps / e2e_s, "unit": "iterations/s", "seconds_per_train_call": e2e_s,
               "iterations_per_call": args.steps,
               "h2d_bytes_per_step": 12 * nnz, "d2h_bytes_per_step": 4 * (nu + ni) * k + (nu + ni),
               "ingest_ms": st2["last_ingest_ms"], "run_ms": st2["last_run_ms"],
               "note": "one step = one pio_als_train call: H2D COO + ingest + K iterations + D2H factors"}
        m2.close()

    if rank != 0:
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------
    # One iteration = item half-step + user half-step; each is (YtY) + one solve launch for the rows up to the
    # heavy-row threshold (+ a part launch and a finish launch for longer rows).  The dominant kernel is the solve
    # launch of the slower half-step; its time is the CUDA-event time of that half-step's solve launches (measured
    # inside pio_als_run on the launching stream), its work the algorithmic FLOPs / bytes of that side.
    hbm_peak, bf16_peak, peak_src = measured_peaks()
    nua = st["n_users_active"] if world == 1 else nu
    nia = st["n_items_active"] if world == 1 else ni
    b_alg, f_solve, f_gram = algorithmic_work(nua, nia, nnz_eff, k, implicit)
    ph = m.phase_ms()
    side_flops = {"user": 2 * nnz_eff * (k * (k + 1) + 2 * k) / 2 + nua * (k ** 3 / 3 + 2 * k * k),
                  "item": 2 * nnz_eff * (k * (k + 1) + 2 * k) / 2 + nia * (k ** 3 / 3 + 2 * k * k)}
    side_bytes = {"user": 8 * nnz_eff + 4 * nua * k + 4 * nia * k, "item": 8 * nnz_eff + 4 * nia * k + 4 * nua * k}
    side_ms = {"user": ph["user_solve_ms"] / args.steps, "item": ph["item_solve_ms"] / args.steps}
    side_kernel = {"user": ph["user_kernel"], "item": ph["item_kernel"]}
    side_tc = {sd: side_kernel[sd] != "fp32" for sd in side_kernel}   # Gramian on tensor cores (tcgen05 or mma.sync)
    kernel_names = {"fp32": "als_solve_kernel (gather + FP32 Gramian + warp Cholesky)",
                    "tcgen05": "tc::als_solve_tc_kernel (tcgen05 split-TF32 Gramian + warp Cholesky)",
                    "mma": "mm::als_solve_mma_kernel (one warp per row: mma.sync 3xTF32 Gramian + warp Cholesky)"}
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    tf32_peak = bf16_peak / 2.0
    split_peak = tf32_peak / 3.0   # an fp32-class product costs three TF32 MMAs (hi*hi + lo*hi + hi*lo)

    def side_obj(sd):
        tf = side_flops[sd] / (side_ms[sd] / 1e3) / 1e12 / max(world, 1)
        gb = side_bytes[sd] / (side_ms[sd] / 1e3) / 1e9 / max(world, 1)
        o = {"kernel": kernel_names[side_kernel[sd]] + f", {sd} half-step",
             "ms_per_launch": side_ms[sd], "algorithmic_flops": side_flops[sd], "algorithmic_bytes": side_bytes[sd],
             "achieved_tflops": tf, "achieved_gbs": gb, "frac_of_fp32_fma_peak": tf / fp32_peak,
             "frac_of_hbm_peak": gb / hbm_peak}
        if side_tc[sd]:
            o["frac_of_split_tf32_tensor_peak"] = tf / split_peak
        return o

    dom = "user" if side_ms["user"] >= side_ms["item"] else "item"
    other = "item" if dom == "user" else "user"
    d = side_obj(dom)
    roofline = {
        "kernel": d["kernel"],
        "bound": "fp32_fma" if not side_tc[dom] else "tensor",
        "achieved": d["achieved_tflops"], "peak": fp32_peak if not side_tc[dom] else split_peak, "unit": "TFLOP/s",
        "frac": d["achieved_tflops"] / (fp32_peak if not side_tc[dom] else split_peak),
        "frac_of_fp32_fma_peak": d["achieved_tflops"] / fp32_peak,
        "peak_source": ("nominal 148 SM x 128 FFMA lanes x 2 x 1.965 GHz (CUDA-core FP32; MEASURED_PEAKS.json has no "
                        "FP32 entry)") if not side_tc[dom] else
                       f"{peak_src} dense bf16 / 2 = tf32 MMA rate, / 3 because an fp32-class product is three TF32 MMAs; "
                       "the kernel is paced by its warp-level Cholesky and fragment loads, not by the tensor pipe",
        "ms_per_launch": d["ms_per_launch"],
        "traffic": traffic_from_profiles(args.workload),
        "traffic_unit": "dram bytes of this launch, from the committed ncu --set full capture (profiles/traffic.json)",
        "per_gpu": True,
        "hbm": {"achieved": d["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": d["achieved_gbs"] / hbm_peak,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": d["algorithmic_bytes"]},
        "other_half_step": side_obj(other),
        "solve_ms_per_iteration": solve_ms / args.steps, "gram_ms_per_iteration": gram_ms / args.steps,
        "comm_ms_per_iteration": comm_ms / args.steps,
        "launches_per_iteration": solve_launches / max(args.steps, 1),
        "algorithmic_flops_per_iteration": f_solve, "algorithmic_bytes_per_iteration": b_alg,
    }
    out = {"metric": "ALS iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
           "wall_ms_per_step": wall_ms / args.steps, "ingest_ms": ingest_ms, "nnz_after_dedup": int(nnz_eff),
           "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu_baseline and args.gpus == 1:
        coo = (du.cpu().numpy(), di.cpu().numpy(), dr.cpu().numpy())
        out["cpu_baseline"] = cpu_baseline_sample(nu, ni, nnz, k, implicit, d_coo=coo)
    print(json.dumps(out))


def nccl_id_2(native, rank, world):
    """A second communicator id for the e2e handle (broadcast from rank 0)."""
    import torch
    import torch.distributed as dist
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(native.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())


if __name__ == "__main__":
    main()
    try:   # leave the torch.distributed group cleanly (no teardown warning after the JSON line)
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
