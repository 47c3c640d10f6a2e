#!/usr/bin/env python
"""bench.py -- DSMIL aggregator forward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

metric  : patches/sec of the DSMIL forward (MILNet.forward, dsmil.py:70-74) at N=10 000, D=512, C=2.
step    : one pass over a stream of `--bags` synthetic bags (default 16 x 10 000 x 512 fp32 = 328 MB,
          larger than the 126 MB L2, so every bag is read from HBM: "inputs larger than L2").
N GPUs  : weak scaling -- every bag is a giant bag of 10 000*N rows row-sharded over the N ranks
          (each rank keeps 10 000 rows per bag); two NCCL all-gathers per step carry the per-class
          critical-instance candidates and the softmax/partial-sum records.
value   : whole-job patches/sec, inputs resident in HBM, CUDA-event timed, max over ranks.
e2e     : same metric through the public host-buffer API (dsmil_wsi_b200.pipeline.HostBagPipeline):
          pinned host bags -> H2D -> forward -> D2H of (classes, prediction_bag, A, B), per step.
roofline: dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json.
cpu_baseline: the oracle's torch-CPU port of the reference op sequence on the host cores (bounded sample).
`--impl reference` times that CPU port alone (the reference is pure PyTorch).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

D, C, NBAG = 512, 2, 10000
METRIC = "patches/sec DSMIL fwd at N=10k D=512"


def algorithmic_bytes_fwd(N, D_, C_):
    """one read of X, write classes + A, weights once, B and pred."""
    W = 4 * (C_ * D_ + C_ + 128 * D_ + 128 + 128 * 128 + 128 + C_ * C_ * D_ + C_)
    return N * (4 * D_ + 8 * C_) + W + 4 * C_ * D_ + 4 * C_


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return float(pk["hbm_gbs"]), float(pk.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def warmup_plan(world: int, warmup: int):
    """(fixed_steps, timed_seconds, extra_fixed_steps).  Under torchrun every step contains collectives, so the
    number of warm-up steps must be the same on all ranks: fixed counts only.  A single process may extend the
    warm-up by wall-clock time so that the clock sampler sees >= 0.6 s of load."""
    fixed = max(int(warmup), 3)
    return (fixed, 0.6, 0) if world == 1 else (fixed, 0.0, 512)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up (0.6 s of identical steps at N=1, 512 steps under torchrun) + timed region, nvidia-smi -lms 100"}


class Weights:
    """Seeded nn.Linear-scale weights for the benchmark model (random init of the reference architecture)."""

    def __init__(self, seed=0, scale=2.0):
        rng = np.random.default_rng(seed)
        u = lambda shape, fan: (rng.uniform(-1, 1, size=shape) * scale / np.sqrt(fan)).astype(np.float32)
        self.Wi, self.bi = u((C, D), D), u((C,), D)
        self.W1, self.b1 = u((128, D), D), u((128,), D)
        self.W2, self.b2 = u((128, 128), 128), u((128,), 128)
        self.Wf, self.bf = u((C, C, D), C * D), u((C,), C * D)


def make_params(seed=0):
    return Weights(seed)


def oracle_params(w):
    """CPU-baseline legs only: hand the same weights to the oracle's torch-CPU port."""
    from oracle import dsmil_oracle as orc
    return orc.Params(w.Wi, w.bi, w.Wf, w.bf, w.W1, w.b1, w.W2, w.b2)


def make_net(p, device):
    import dsmil as mil
    net = mil.MILNet(mil.FCLayer(D, C), mil.BClassifier(D, C))
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float32))
    net.load_state_dict({"i_classifier.fc.0.weight": t(p.Wi), "i_classifier.fc.0.bias": t(p.bi),
                         "b_classifier.q.0.weight": t(p.W1), "b_classifier.q.0.bias": t(p.b1),
                         "b_classifier.q.2.weight": t(p.W2), "b_classifier.q.2.bias": t(p.b2),
                         "b_classifier.fcc.weight": t(p.Wf), "b_classifier.fcc.bias": t(p.bf)})
    return net.to(device).eval()


def best_cpu_threads(p, bags, budget=0.6):
    """torch-CPU with one thread per core is NOT the fastest setting on a many-core host for ops this
    small; give the baseline the thread count it likes best (short calibration, reported as `cores`)."""
    from oracle import dsmil_oracle as orc
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_rate = cands[0], 0.0
    for c in cands:
        port = orc.TorchPort(oracle_params(p), threads=c)
        port.forward(bags[0])
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < budget:
            port.forward(bags[n % len(bags)]); n += 1
        rate = n / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = c, rate
    return best


def cpu_port_rate(p, seconds, threads=None, nbags=4):
    """Oracle torch-CPU port (reference op sequence) on a bounded sample: returns patches/s."""
    from oracle import dsmil_oracle as orc
    g = torch.Generator().manual_seed(1)
    bags = [torch.rand(NBAG, D, generator=g) for _ in range(nbags)]
    threads = threads or best_cpu_threads(p, bags)
    port = orc.TorchPort(oracle_params(p), threads=threads)
    for b in bags[:2]:
        port.forward(b)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        port.forward(bags[n % nbags])
        n += 1
    dt = time.perf_counter() - t0
    return n * NBAG / dt, port.threads, n


def run_reference(args):
    """`--impl reference`: the reference is pure PyTorch, so its CPU implementation of the path is the
    torch-CPU op sequence of dsmil.py; timed here via the oracle port (no /root/reference on the box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import dsmil_oracle as orc
    g = torch.Generator().manual_seed(1)
    nb = args.ref_bags
    bags = [torch.rand(NBAG, D, generator=g) for _ in range(nb)]
    port = orc.TorchPort(oracle_params(make_params()), threads=best_cpu_threads(make_params(), bags[:4]))
    for _ in range(max(args.warmup, 1)):
        for b in bags:
            port.forward(b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for b in bags:
            port.forward(b)
    dt = time.perf_counter() - t0
    val = args.steps * nb * NBAG / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "patches/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"stream of {nb} synthetic bags, each N={NBAG} x D={D}, C={C}, DSMIL forward "
                                  "(bounded sample of the GPU arm's 16-bag step)", "bags_per_step": nb},
           "cpu_baseline": {"value": val, "unit": "patches/s", "cores": port.threads, "kind": "port",
                            "sample": f"{args.steps} steps x {nb} bags x {NBAG} patches, torch-CPU fp32, all host threads"},
           "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the DSMIL B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a collective that cannot complete (e.g. a rank died) aborts after 3 minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from dsmil_wsi_b200 import _lib
    from dsmil_wsi_b200.pipeline import HostBagPipeline
    from dsmil_wsi_b200.sharded import (CudaShardBagOps, CudaShardOps, milnet_params, sharded_forward_bags,
                                        sharded_forward_bags_batched)
    lib = _lib.load()

    p = make_params()
    net = make_net(p, dev)
    nb = args.bags
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    bags = [torch.rand(NBAG, D, generator=g, device=dev) for _ in range(nb)]   # this rank's rows of each bag
    offsets = [rank * NBAG] * nb
    ops = bops = None
    if world > 1:
        if CudaShardBagOps.supported(milnet_params(net)):
            bops = CudaShardBagOps(milnet_params(net))
        else:
            ops = CudaShardOps(milnet_params(net))

    def sharded_step(xs):
        if bops is not None:
            return sharded_forward_bags_batched(bops, xs, offsets)
        return sharded_forward_bags(ops, xs, offsets)

    def step():
        with torch.no_grad():
            if world == 1:
                return net.forward_bags(bags)
            return sharded_step(bags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up on (the timed region alone lasts a few ms, shorter
    # than nvidia-smi's fastest period).  The number of warm-up steps MUST be identical on every rank (each
    # step contains collectives), so it is a fixed count under torchrun and time-based only for a single process.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fixed_steps, timed_s, extra_steps = warmup_plan(world, args.warmup)
    for _ in range(fixed_steps):
        step()
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    n_w = 0
    while time.perf_counter() - t_w < timed_s:        # single process only (timed_s == 0 under torchrun)
        step()
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize()
    for n_w in range(1, extra_steps + 1):              # same count on all ranks
        step()
        if n_w % 8 == 0:
            torch.cuda.synchronize()
    barrier()
    l0 = lib.dsmil_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = int(lib.dsmil_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    patches_per_step = nb * NBAG * world
    value = patches_per_step / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launch stream -------
    hbm_peak, tf_peak, peak_src = load_peaks()
    lib.dsmil_profile_enable(1)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ms_tag = (ctypes.c_double * 8)()
    n_tag = (ctypes.c_uint64 * 8)()
    lib.dsmil_profile_read(ms_tag, n_tag)
    lib.dsmil_profile_enable(0)
    tags = ["scores", "q_mlp", "attend", "finalize", "fused_sm100"]
    per = {t: (ms_tag[i] / n_tag[i] if n_tag[i] else None) for i, t in enumerate(tags)}
    dom = max((t for t in tags if per[t]), key=lambda t: per[t] * n_tag[tags.index(t)])
    launches_dom = int(n_tag[tags.index(dom)])
    bags_per_launch = 2.0 * nb / launches_dom          # 2 profiled steps of nb bags each
    alg = algorithmic_bytes_fwd(NBAG, D, C) * bags_per_launch
    dom_ms = per[dom]
    achieved = alg / (dom_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):     # dram__bytes_read+write of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(tpath))
        if tj.get("tag") == dom:
            traffic = tj["dram_bytes_per_row"] * NBAG * bags_per_launch
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "bags_per_launch": bags_per_launch, "kernel_ms": dom_ms,
                "per_kernel_ms": {k: v for k, v in per.items() if v},
                "whole_forward_frac": (algorithmic_bytes_fwd(NBAG, D, C) * nb / (ms_per_step / 1e3) / 1e9) / hbm_peak,
                "note": "achieved = algorithmic bytes of the whole fused forward (SURVEY 8d: 2048+8C B/patch + weights) for the "
                        "bags one launch covers / the dominant kernel's mean CUDA-event duration; the forward is three "
                        "kernels (phase 1 tcgen05 Q-MLP+scores, attend, finalize): whole_forward_frac charges all of them "
                        "(CUDA-event time of the full step) and is the honest end-to-end roofline fraction"}

    # ---- end to end through the public host-buffer API ------------------------------------------
    e2e = None
    if world == 1:
        host = [b.cpu().pin_memory() for b in bags]
        pipe = HostBagPipeline(net, NBAG, D, C)
        for _ in range(2):
            pipe.run(host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 10))
        for _ in range(reps):
            pipe.run(host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        h2d, d2h = pipe.bytes_per_bag(NBAG)
        e2e = {"value": nb * NBAG / dt, "unit": "patches/s", "h2d_bytes_per_step": h2d * nb,
               "d2h_bytes_per_step": d2h * nb, "ms_per_step": dt * 1e3,
               "api": "dsmil_wsi_b200.pipeline.HostBagPipeline.run(pinned host bags)"}
    else:
        # multi-GPU e2e: every rank stages its shard from pinned host memory, then the sharded forward
        host = [b.cpu().pin_memory() for b in bags]
        slots = [torch.empty_like(b) for b in bags]
        outs_host = None

        def e2e_step():
            with torch.no_grad():
                for s, h in zip(slots, host):
                    s.copy_(h, non_blocking=True)
                outs = sharded_step(slots)
                return [tuple(t.cpu() for t in o[:4]) for o in outs]
        e2e_step(); barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": patches_per_step / dt, "unit": "patches/s", "h2d_bytes_per_step": 4 * NBAG * D * nb * world,
               "d2h_bytes_per_step": 4 * (2 * NBAG * C + C + C * D) * nb * world, "ms_per_step": dt * 1e3,
               "api": "pinned host shards -> sharded_forward_bags -> host"}

    cpu_baseline = None
    if rank == 0 and world == 1:
        rate, threads, nrun = cpu_port_rate(p, args.cpu_seconds)
        cpu_baseline = {"value": rate, "unit": "patches/s", "cores": threads, "kind": "port",
                        "sample": f"{nrun} forwards of one N={NBAG} bag in ~{args.cpu_seconds:.0f}s, torch-CPU fp32 "
                                  "port of dsmil.py:46-62 (oracle/dsmil_oracle.py TorchPort)"}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "slides_per_sec": value / (NBAG * world),
               "config": {"workload": f"stream of {nb} synthetic bags per step, each N={NBAG * world} x D={D} fp32 "
                                      f"(U[0,1)), C={C}, nonlinear q, identity v; MILNet forward; "
                                      f"{'one GPU' if world == 1 else f'rows sharded over {world} GPUs ({NBAG} rows/rank/bag)'}",
                          "bags_per_step": nb, "rows_per_rank_per_bag": NBAG, "parallelism": f"row-shard x{world}",
                          "l2_policy": f"inputs larger than L2: {nb} bags x {NBAG * D * 4 / 1e6:.1f} MB per rank cycled",
                          "forward_path": int(lib.dsmil_forward_path(ctypes.byref(_lib.DsmilParams(D, C, 1, 0)), NBAG))},
               "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches * world,
               "clocks": clocks}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bags", type=int, default=16, help="bags per step (16 x 20.5 MB > L2)")
    ap.add_argument("--ref-bags", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
