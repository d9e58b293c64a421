#!/usr/bin/env python
"""bench.py - TP-GAN G+D training-step throughput (images/sec) on N B200s, and the reference arm on the host CPU.

Own arm (default):  one process per GPU; per-GPU batch 32 (BASELINE.json configs[1]: "full G+D training step batch 32 on
1xB200 (tf32)"); synthetic 128x128 faces + landmark patches (oracle/step.py conventions); weights = seeded
reference init.  A step = G forward, D phase (WGAN-GP critic incl. gradient penalty, Adam), G phase (all losses, backward,
Adam, weight repack) - nothing is skipped.  `value` = images/s with the batch already resident in HBM; `e2e` = the same
through the public step() call with HOST (pinned) inputs copied in and the metrics read back every step.
Reference arm (--impl reference): the oracle's fp32 PyTorch restatement of the reference modules + step (the reference
itself, /root/reference, does not travel to the GPU box and ships no training step), on all host cores.
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

METRIC = "tpgan_g_d_train_step_images_per_sec"
UNIT = "images/s"
PER_GPU_BATCH = 32
CPU_BATCH = 4          # BASELINE.json configs[0]: the reference's own CPU-runnable case


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.proc = None
        self.index = index
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
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
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step_time(batch: int, steps: int, warmup: int):
    """Seconds per oracle-port training step (fp32 PyTorch, all host threads) at `batch` images."""
    import torch
    from oracle import model_port as mp, step as ostep
    from tpgan_b200 import D_and_G_model as M, config
    torch.manual_seed(0)
    G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"])
    D = M.Discriminator(config.D["use_batchnorm"])
    pg = {k: v.clone().requires_grad_(True) for k, v in G.state_dict().items()}
    pd = {k: v.clone().requires_grad_(True) for k, v in D.state_dict().items()}
    del G, D
    Gc, Dc = ostep.port_callables(pg, pd)
    og = torch.optim.Adam(list(pg.values()), lr=ostep.LEARNING_RATE)
    od = torch.optim.Adam(list(pd.values()), lr=ostep.LEARNING_RATE)
    b = ostep.make_batch(batch)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ostep.train_step(Gc, Dc, list(pg.values()), list(pd.values()), og, od, b)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = CPU_BATCH      # BASELINE.json configs[0] / BASELINE.md section 4: batch 4, >= 1 warm-up step, >= 3 timed steps
    steps, warmup = max(a.steps, 3), max(a.warmup, 1)
    sec, threads = cpu_port_step_time(batch, steps, warmup)
    val = batch / sec
    sample = (f"oracle fp32 PyTorch port of the reference G+D + oracle step, batch {batch} per step, {steps} timed steps after "
              f"{warmup} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TP-GAN G+D training step (WGAN-GP critic + all G losses + Adam), 128x128, "
                                   f"bounded CPU sample batch {batch}", "per_step_batch": batch},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _ncu_summary(dtype: str, batch: int):
    """Per-kernel DRAM traffic and tensor-pipe activity from the committed ncu pass of this same workload
    (profiles/ncu_kernel_summary_r2.json, produced by tools/ncu_summary.py from `ncu --metrics gpu__time_duration.sum,
    dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...` over one eager step)."""
    p = os.path.join(ROOT, "profiles", "ncu_kernel_summary_r2.json")
    if not os.path.exists(p):
        return {}
    d = json.load(open(p))
    return d.get(f"{dtype}_b{batch}", {})


def gan_line(a, B, dtype, identity, steps, warmup, world, rank, dev, want_roofline=True, want_cpu=True, per_layer=""):
    """Build a trainer, time `steps` resident steps and `steps` end-to-end steps, optionally the per-kernel roofline.
    Returns the JSON line (rank 0) or None."""
    import gc

    import torch
    import torch.distributed as dist
    from tpgan_b200 import D_and_G_model as M, _lib, config, synthetic
    from tpgan_b200.train_step import Eager, TPGANTrainer
    local = dev.index
    torch.manual_seed(0)
    G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"])
    D = M.Discriminator(config.D["use_batchnorm"])
    G.to(dev)
    D.to(dev)
    ident = None
    if identity:   # BASELINE configs[2]: frozen FeatureExtract (ResNet18-128) identity-preserving loss
        from tpgan_b200.FeatureExtract import FeatureExtractModel
        from tpgan_b200.ResNet import BasicBlock
        ident = FeatureExtractModel("resnet", config.G["num_classes"], residualBlock=BasicBlock,
                                    feature_layer_dim_before_FC=256).to(dev).eval()
    if a.sm_reserve:
        _lib.set_sm_reserve(a.sm_reserve)
    tr = TPGANTrainer(G, D, B, device=dev, use_dropout=True, world_size=world, use_graphs=not a.no_graphs,
                      identity_net=ident, dtype=dtype, overlap_allreduce=a.dp_mode == "overlap",
                      graph_collectives=a.graph_collectives, bucket_mb=a.bucket_mb)
    host = synthetic.make_batch(B, seed=1234 + rank)
    host = {k: host[k].contiguous().pin_memory() for k in synthetic.KEYS}
    devb = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    resident = lambda: tr.step(devb, read_metrics=False)
    d2h_bytes = [0]

    def e2e_step():
        # pinned host tensors in, metrics read back to the host, every step.  The H2D copies of the NEXT step's batch are
        # started on a copy stream inside this call (TPGANTrainer.prefetch) and overlap this step's kernels - one H2D of a
        # full batch and one D2H of the metrics per step, as a prefetching data loader would drive the public API.
        m = tr.step(host, read_metrics=True, prefetch_next=host)
        d2h_bytes[0] = 16 * 4 + 3 * B * 16 * 4 * 4 + 4
        return m

    for _ in range(max(warmup, 3)):
        resident()
    if getattr(a, "profile_step", False):
        # ncu --profile-from-start off: exactly ONE step between cudaProfilerStart/Stop (use with --no-graphs so that every
        # kernel is its own launch); numbers printed by a run under ncu are never bench values
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(resident, steps)
    launches = _lib.launch_count() - l0
    if not a.no_graphs:   # replayed graphs do not go through the library's host entry points: count the captured kernels
        runner = list(tr._sched.values())[0]
        launches = steps * runner.kernels_per_run
    clocks = sampler.stop() if rank == 0 else None
    tr.prefetch(host)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    assert _lib.kernel_status() == 0, "a kernel aborted a barrier wait"
    # replicas must stay identical under data parallelism: spread of a parameter checksum over the ranks
    spread = None
    if world > 1:
        cs = torch.stack([tr.flat_g.data.double().sum(), tr.flat_d.data.double().sum()])
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        st = torch.stack(allcs)
        spread = float(((st.max(0).values - st.min(0).values) / st.mean(0).abs().clamp_min(1e-30)).max())

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv): per-launch CUDA events over one more (eager) step of
    # the WHOLE schedule - generator plan, critic (D phase, gradient penalty tangent pass, G phase), identity network
    roof = None
    cpu = None
    if rank == 0 and want_roofline:
        ev = []
        torch.cuda.synchronize()
        tr.load_inputs(devb)
        # eager launches cost the host ~10-20 us each (planning + tensor-map encoding), more than the small kernels run: keep
        # the device queue full (the host enqueues the whole step behind a ~30 ms spin) so that the events bracket kernel
        # execution, not the GPU waiting for the next launch
        torch.cuda._sleep(int(0.03 * 1.9e9))
        for f in tr._schedule(True):
            if isinstance(f, Eager):
                continue                 # collectives: the other ranks are not here
            if getattr(f, "kind", None) in ("tapgemm", "rowconv", "wgrad"):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                f()
                e.record()
                # conv launches: ask the library which kernel it actually routed the call to
                kind = f.kind if f.kind == "wgrad" else _lib.last_conv_kernel()
                ev.append((kind, f.flops, s, e, f.label))
            else:
                f()
        torch.cuda.synchronize()
        agg = {}
        for kind, fl, s, e, label in ev:
            t = s.elapsed_time(e)
            g = agg.setdefault(kind, [0.0, 0.0, 0])
            g[0] += fl
            g[1] += t
            g[2] += 1
        if per_layer:
            rows = sorted(((s.elapsed_time(e), kind, fl, label) for kind, fl, s, e, label in ev), reverse=True)
            with open(per_layer, "w") as f:
                for t, kind, fl, label in rows:
                    f.write(json.dumps({"ms": round(t, 4), "kind": kind, "tflops": round(fl / (t * 1e-3) / 1e12, 1),
                                        "gflop": round(fl / 1e9, 2), "label": label}) + "\n")
        peaks, which = _peaks()
        bf16 = dtype == "bf16"
        div = 1.0 if bf16 else 2.0
        peak_sus = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / div
        peak_burst = peaks["bf16_tflops"] / div
        names = {"tapgemm": f"tapgemm_kernel (tcgen05 {dtype} multi-tap implicit-GEMM conv fwd/dgrad/deconv/linear)",
                 "rowconv": f"rowconv_kernel (tcgen05 {dtype} row-tile conv for the wide 128x128 stride-1 layers, fwd/dgrad)",
                 "rowstack": f"rowstack_kernel (tcgen05 {dtype} N-stacked row-tile conv for the narrow 128x128 layers, fwd/dgrad)",
                 "flatconv": f"flatconv_kernel (tcgen05 {dtype} flat-slab conv for the small-map stride-1 layers, fwd/dgrad)",
                 "wgrad": f"wgrad_kernel (tcgen05 {dtype} weight-gradient GEMM over pixels)"}
        step_ms = ms / steps
        ncu = _ncu_summary(dtype, B)

        def entry(kind):
            fl, t, n = agg[kind]
            ach = fl / (t * 1e-3) / 1e12
            d = {"kernel": names[kind], "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus,
                 "frac_of_burst_peak": ach / peak_burst, "launches": n, "avg_launch_ms": t / n, "share_of_step": t / step_ms,
                 "algorithmic_gflop_per_step": fl / 1e9}
            k = ncu.get(kind + "_kernel")
            if k:     # from the committed ncu pass of this workload (per launch, like `achieved`)
                d.update(traffic=k["dram_bytes_per_launch"], tensor_pipe_pct=k["tensor_pipe_pct"],
                         ncu_share_of_step=k["share_of_step"])
            return d
        dom = max(agg, key=lambda k: agg[k][1])
        tot_fl, tot_t = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
        roof = {"bound": "tensor", "traffic": None, **entry(dom),
                "traffic_source": ("profiles/ncu_kernel_summary_r2.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                   "averaged over the kernel's launches of one step") if ncu else None,
                "peak_source": (f"bf16_tflops_sustained of MEASURED_PEAKS.json ({which}); frac_of_burst_peak uses bf16_tflops" if bf16
                                else f"0.5 x bf16_tflops_sustained of MEASURED_PEAKS.json ({which}); tf32 = half the bf16 rate "
                                     "(assumed, not measured); frac_of_burst_peak uses 0.5 x bf16_tflops"),
                "scope": "every tensor-core launch of the step: generator plan, critic D / GP-tangent / G phases" +
                         (", identity network" if identity else ""),
                "other_kernels": {k: entry(k) for k in agg if k != dom},
                "all_conv_kernels": {"achieved": tot_fl / (tot_t * 1e-3) / 1e12, "frac": tot_fl / (tot_t * 1e-3) / 1e12 / peak_sus,
                                     "share_of_step": tot_t / step_ms, "algorithmic_gflop_per_step": tot_fl / 1e9},
                "whole_step": {"achieved": tot_fl / (step_ms * 1e-3) / 1e12, "frac": tot_fl / (step_ms * 1e-3) / 1e12 / peak_sus}}
    if rank == 0 and world == 1 and want_cpu and not a.no_cpu:   # N = 1 only: with N ranks the host cores are shared
        sec, threads = cpu_port_step_time(CPU_BATCH, 3, 1)
        cpu = {"value": CPU_BATCH / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"three oracle-port G+D training steps (fp32 PyTorch, CPU) at batch {CPU_BATCH} after one warm-up step"}
    line = None
    if rank == 0:
        gb = B * world
        line = {"metric": METRIC, "value": gb * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": max(warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "strong" if a.global_batch else "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": f"TP-GAN G+D training step (WGAN-GP critic + all G losses + Adam), batch {B}/GPU, "
                                       "128x128 synthetic faces + 4 landmark patches, dropout on" +
                                       (", + frozen ResNet18 identity-preserving loss" if identity else ""),
                           "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{world}",
                           "l2": f"activations per step (~{0.27 * B:.0f} GB) exceed the 126 MB L2; no explicit flush",
                           "cuda_graphs": not a.no_graphs,
                           **({"dp_mode": a.dp_mode, "graph_collectives": a.graph_collectives, "sm_reserve": a.sm_reserve,
                               "bucket_mb": a.bucket_mb} if world > 1 else {})},
                "clocks": clocks,
                "e2e": {"value": gb * steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "h2d": "pinned host batch, copied on a side stream during the previous step",
                        "d2h_bytes_per_step": d2h_bytes[0], "ms_per_step": ms_e2e / steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        if spread is not None:
            line["replica_checksum_spread"] = spread
    del tr, G, D, ident, devb
    gc.collect()
    torch.cuda.empty_cache()
    return line


def run_b200(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = a.batch
    if a.global_batch:
        assert a.global_batch % world == 0, "--global-batch must be a multiple of the number of GPUs"
        B = a.global_batch // world
    line = gan_line(a, B, a.dtype, a.identity, a.steps, a.warmup, world, rank, dev, per_layer=a.per_layer)
    if rank == 0 and world == 1 and not a.no_secondary and not a.identity and a.dtype == "tf32" and not a.global_batch:
        # the other single-GPU configurations BASELINE.json names, measured in the same run (short: 5 timed steps each)
        sec = []
        try:
            l2 = gan_line(a, 64, "bf16", True, 5, 3, 1, 0, dev, want_roofline=False, want_cpu=False)
            sec.append({"config": "BASELINE configs[2]: G+D step + frozen ResNet18 identity loss, bf16 operands, batch 64, 1 GPU",
                        **{k: l2[k] for k in ("value", "unit", "ms_per_step", "dtype", "steps", "e2e", "gpu_launches", "clocks")}})
            l3 = gan_line(a, 32, "bf16", False, 5, 3, 1, 0, dev, want_roofline=True, want_cpu=False)
            sec.append({"config": "the headline workload (configs[1], batch 32) with bf16 operands",
                        **{k: l3[k] for k in ("value", "unit", "ms_per_step", "dtype", "steps", "e2e", "gpu_launches")},
                        "roofline": {k: l3["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "share_of_step",
                                                                    "all_conv_kernels", "whole_step")}})
            for backbone in ("mobilenetv2", "resnet"):
                ns = argparse.Namespace(**{**vars(a), "batch": 32, "steps": 10, "warmup": 3, "no_cpu": True, "backbone": backbone})
                lp = pretrain_line(ns, 1, 0, dev, want_roofline=False)
                sec.append({"config": f"BASELINE configs[4] per GPU: Pretrain step, {backbone} backbone, batch 32, 1 GPU",
                            **{k: lp[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "steps", "e2e", "gpu_launches")},
                            **({"single_pass_tf32": lp["single_pass_tf32"]} if "single_pass_tf32" in lp else {})})
        except Exception as ex:     # a secondary measurement must never take the headline down with it
            sec.append({"error": f"{type(ex).__name__}: {ex}"})
        line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ Pretrain workload
PRETRAIN_METRIC = "mobilenetv2_pretrain_step_images_per_sec"
PRETRAIN_METRIC_RESNET = "resnet18_pretrain_step_images_per_sec"


def cpu_pretrain_step_time(batch: int, steps: int, warmup: int):
    """Seconds per oracle-port Pretrain step (MobileNetV2 + MultiTaskLoss + SGD-Nesterov, fp32 PyTorch, all host threads)."""
    import torch
    from oracle import pretrain_port as P
    torch.manual_seed(0)
    net = P.MobileNetV2Port()
    opt = torch.optim.SGD(net.parameters(), **P.SGD)
    x, true, u = P.make_batch(batch)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        P.pretrain_step(net, x, true, u, opt)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), torch.get_num_threads()


def cpu_classifier_step_time(batch: int, steps: int, warmup: int):
    """Seconds per oracle-port ResNet18-128 classifier pre-training step (fp32 PyTorch, all host threads)."""
    import torch
    import torch.nn.functional as F
    from oracle import identity_port as ip
    from tpgan_b200.ResNet import BasicBlock, ResNet18
    torch.manual_seed(0)
    net = ResNet18(BasicBlock, 347, True, 256)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    params = [sd[k].requires_grad_(True) for k in names]
    from oracle.pretrain_port import SGD
    opt = torch.optim.SGD(params, **SGD)
    g = torch.Generator().manual_seed(1)
    x, y = torch.rand((batch, 3, 128, 128), generator=g) * 2 - 1, torch.randint(0, 347, (batch,), generator=g)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        F.cross_entropy(ip.resnet18_128(sd, x, training=True)[0], y).backward()
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), torch.get_num_threads()


def run_pretrain_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    batch = a.batch
    resnet = a.backbone == "resnet"
    sec, threads = (cpu_classifier_step_time if resnet else cpu_pretrain_step_time)(batch, a.steps, max(a.warmup, 1))
    val = batch / sec
    sample = (f"oracle fp32 PyTorch port of ResNet18-128 + cross-entropy + SGD step, batch {batch} per step, {a.steps} steps"
              if resnet else
              f"oracle fp32 PyTorch port of MobileNetV2 + MultiTaskLoss + SGD step, batch {batch} per step, {a.steps} steps")
    line = {"impl": "reference", "metric": PRETRAIN_METRIC_RESNET if resnet else PRETRAIN_METRIC, "value": val, "unit": UNIT,
            "n_gpus": a.gpus, "steps": a.steps,
            "warmup": max(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"feature-extractor pre-training step: ResNet18-128 + cross-entropy + SGD-Nesterov, batch {batch}"
                                    if resnet else
                                    f"Pretrain.py step: MobileNetV2-SSD + MultiTaskLoss + SGD-Nesterov, batch {batch}, 128x128"),
                       "per_step_batch": batch},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_pretrain(a):
    """BASELINE.json configs[4]: Pretrain.py feature-extractor pre-training, MobileNetV2, global batch 256 at 8 GPUs
    (32 per GPU, weak scaling), synthetic 128x128 faces."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = pretrain_line(a, world, rank, dev)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _pretrain_batch(B: int, seed: int):
    """Synthetic Pretrain batch: faces U(-1,1), 4 ground-truth points = landmark means + U(-3,3) px jitter,
    per-point sub-sampling keys.  Same draw order as the oracle's generator (product code: no oracle import)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((B, 3, 128, 128), generator=g) * 2 - 1
    # eyes, nose, mouth centre (mean of the two mouth corners) of D_and_G_model.py:120-128
    means = torch.tensor([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [64.7803, 89.3250]])
    true = (means[None] + (torch.rand((B, 4, 2), generator=g) * 6 - 3)).reshape(B, 8)
    u = torch.rand((B, 394), generator=g)
    return x, true, u


def pretrain_line(a, world, rank, dev, want_roofline=True):
    import torch
    import torch.distributed as dist
    local = dev.index
    from tpgan_b200 import _lib
    from tpgan_b200.MobileNetV2 import MobileNetV2
    from tpgan_b200.pretrain_step import ClassifierTrainer, PretrainTrainer
    B = a.batch
    resnet = a.backbone == "resnet"
    torch.manual_seed(0)
    x, true, u = _pretrain_batch(B, seed=1234 + rank)
    if resnet:
        from tpgan_b200.FeatureExtract import FeatureExtractModel
        from tpgan_b200.ResNet import BasicBlock
        net = FeatureExtractModel("resnet", 347, residualBlock=BasicBlock, feature_layer_dim_before_FC=256).to(dev)
        tr = ClassifierTrainer(net, B, device=dev, world_size=world, use_graphs=not a.no_graphs)
        host = [x.contiguous().pin_memory(), torch.randint(0, 347, (B,), generator=torch.Generator().manual_seed(rank)).pin_memory()]
    else:
        net = MobileNetV2().to(dev)
        tr = PretrainTrainer(net, B, device=dev, world_size=world, use_graphs=not a.no_graphs)
        host = [t.contiguous().pin_memory() for t in (x, true, u)]
    devb = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    resident = lambda: tr.step(*devb, read_metrics=False)
    hb = tuple(host)
    # pinned host batch in (its copies are started on a copy stream during the previous step), loss read back, every step
    e2e_step = lambda: tr.step(None, None, batch=hb, prefetch_next=hb, read_metrics=True)
    for _ in range(max(a.warmup, 3)):
        resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(resident, a.steps)
    launches = _lib.launch_count() - l0
    if not a.no_graphs:
        launches = a.steps * list(tr._sched.values())[0].kernels_per_run
    clocks = sampler.stop() if rank == 0 else None
    tr.prefetch(hb)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    assert _lib.kernel_status() == 0, "a kernel aborted a barrier wait"
    single = None
    if not resnet:
        # The production mode of this network is the fp32-accurate 3xTF32 operand split (within 1e-3 of the reference in
        # training mode; single-pass tf32 is 3.7e-2 off after ~50 BatchNorm-renormalised layers).  It costs launches - one
        # split + three tensor-core launches per convolution in a launch-bound step - so the single-pass figure is reported
        # next to it (same model, same batch, PretrainTrainer(exact=False)).
        net1 = MobileNetV2().to(dev)
        tr1 = PretrainTrainer(net1, B, device=dev, world_size=world, use_graphs=not a.no_graphs, exact=False)
        r1 = lambda: tr1.step(*devb, read_metrics=False)
        for _ in range(max(a.warmup, 3)):
            r1()
        ms1 = timed(r1, a.steps)
        single = {"value": B * world * a.steps / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / a.steps, "dtype": "tf32",
                  "note": "single-pass tf32 convolutions (PretrainTrainer(exact=False)): training-mode outputs 3.7e-2 from the fp32 oracle"}
        del tr1, net1
    roof = cpu = None
    if rank == 0 and want_roofline:
        ev = []
        tr.load_inputs(*devb)
        tr._stage()
        torch.cuda.synchronize()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2: each timed launch starts cold
        for lst in (tr.plan.fwd, tr.plan.bwd):
            for f in lst:
                kind = getattr(f, "kind", None)
                if kind is None:
                    f()
                    continue
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                f()
                e.record()
                ev.append((kind, getattr(f, "bytes", 0.0), getattr(f, "flops", 0.0), s, e))
        torch.cuda.synchronize()
        agg = {}
        for kind, by, fl, s, e in ev:
            g = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
            g[0] += by
            g[1] += fl
            g[2] += s.elapsed_time(e)
            g[3] += 1
        peaks, which = _peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tf32_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0
        total_t = sum(v[2] for v in agg.values())

        def entry(kind):
            by, fl, t, n = agg[kind]
            d = {"launches": n, "avg_launch_ms": t / n, "share_of_timed_launches": t / total_t}
            if by > 0:
                d.update(achieved=by / (t * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s", frac=by / (t * 1e-3) / 1e9 / hbm_peak,
                         algorithmic_mb_per_step=by / 1e6)
            else:
                d.update(achieved=fl / (t * 1e-3) / 1e12, peak=tf32_peak, unit="TFLOP/s", frac=fl / (t * 1e-3) / 1e12 / tf32_peak,
                         algorithmic_gflop_per_step=fl / 1e9)
            return d
        dom = max(agg, key=lambda k: agg[k][2])
        names = {"bn_fwd": "bn_stats/bn_finalize/bn_apply_kernel (training BatchNorm + ReLU6 / residual forward)",
                 "bn_bwd": "bn_bwd_reduce/bn_bwd_apply_kernel (BatchNorm + ReLU6 backward)",
                 "dw_fwd": "dw3x3_fwd_kernel", "dw_dgrad": "dw3x3_dgrad_kernel", "dw_wgrad": "dw3x3_wgrad_kernel",
                 "tapgemm": "tapgemm_kernel (tcgen05 tf32: 1x1 / 3x3 dense convs fwd + dgrad)",
                 "wgrad": "wgrad_kernel (tcgen05 tf32 weight gradients)"}
        roof = {"bound": "hbm" if agg[dom][0] > 0 else "tensor", "kernel": names.get(dom, dom), **entry(dom), "traffic": None,
                "timing": "per-launch CUDA events in eager mode with an L2 flush (256 MB memset) before each launch",
                "peak_source": f"MEASURED_PEAKS.json ({which})",
                "other_kernels": {names.get(k, k): entry(k) for k in agg if k != dom}}
    line = None
    if rank == 0:
        if not a.no_cpu and world == 1:
            sec, threads = (cpu_classifier_step_time if resnet else cpu_pretrain_step_time)(B, 2, 1)
            cpu = {"value": B / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"two oracle-port {'ResNet18 classifier' if resnet else 'Pretrain'} steps (fp32 PyTorch, CPU) at batch {B} after one warm-up"}
        gb = B * world
        line = {"metric": PRETRAIN_METRIC_RESNET if resnet else PRETRAIN_METRIC, "value": gb * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if resnet else "tf32x3 (fp32-accurate three-pass operand split, the MobileNetV2 production mode)",
                "data": "synthetic",
                "config": {"workload": (f"feature-extractor pre-training step: ResNet18-128 forward/backward (training BatchNorm) + "
                                        f"softmax cross-entropy (347 identities) + SGD-Nesterov, batch {B}/GPU, 128x128 synthetic faces"
                                        if resnet else
                                        f"Pretrain.py step: MobileNetV2-SSD forward/backward (training BatchNorm) + batched "
                                        f"MultiTaskLoss + SGD-Nesterov, batch {B}/GPU, 128x128 synthetic faces"),
                           "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{world}",
                           "l2": "the step's working set (~1.2 GB at batch 32) exceeds the 126 MB L2; no flush between steps",
                           "cuda_graphs": not a.no_graphs},
                "clocks": clocks,
                "e2e": {"value": gb * a.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / a.steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        if single is not None:
            line["single_pass_tf32"] = single
    del tr, net, devb
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    ap.add_argument("--workload", default="gan", choices=["gan", "pretrain"],
                    help="gan = G+D training step (BASELINE configs[1], the headline); pretrain = Pretrain.py MobileNetV2 "
                         "step (configs[4])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true",
                    help="gan workload, 1 GPU: skip the short measurements of BASELINE configs[2] / [4] appended to the line")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="gan workload: fixed GLOBAL batch split over the GPUs (BASELINE configs[3]: 256 at 2/4/8 GPUs = "
                         "128/64/32 per GPU; strong scaling) instead of a fixed per-GPU batch")
    ap.add_argument("--per-layer", default="", help="write the per-launch conv/wgrad timing table (JSON lines) here")
    ap.add_argument("--backbone", default="mobilenetv2", choices=["mobilenetv2", "resnet"],
                    help="pretrain workload: MobileNetV2-SSD landmark pre-training (Pretrain.py) or ResNet18 identity classifier")
    ap.add_argument("--identity", action="store_true", help="gan workload: add the frozen identity network's loss (configs[2], tf32)")
    ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"],
                    help="gan workload: tensor-core operand type (tf32 = BASELINE configs[1]; bf16 = configs[2])")
    ap.add_argument("--dp-mode", default="overlap", choices=["overlap", "serial"],
                    help="N > 1: bucketed all-reduce overlapped with backward, or one all-reduce after backward")
    ap.add_argument("--graph-collectives", action="store_true", help="N > 1: capture the NCCL calls into the step's CUDA graph")
    ap.add_argument("--sm-reserve", type=int, default=0, help="SMs the persistent kernels leave free (for the NCCL kernels)")
    ap.add_argument("--bucket-mb", type=float, default=128.0, help="N > 1, overlap mode: all-reduce bucket size")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket one extra step with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    a = ap.parse_args()
    if a.workload == "pretrain":
        (run_pretrain_reference if a.impl == "reference" else run_pretrain)(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
