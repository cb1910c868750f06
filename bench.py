"""bench.py — SPARC + global InfoNCE fwd+bwd pairs/s (BASELINE.json metric) and AdamSPD step GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

A "step" is one pass of the hot path over one batch of synthetic CLIP-shaped embeddings:
SPARCLoss forward (fine-grained + global InfoNCE) and backward to dv, dl — through the package's
public API, which calls the sm_100a kernels over the C ABI.  Workload = BASELINE config 2
(ViT-B/16: B=256/GPU, P=196, T=77, D=512, bf16, thr=1/P, s=1, all-True mask, seed 42+rank); with
N > 1 the global InfoNCE all-gathers the pooled embeddings over NCCL (weak scaling, fixed B/GPU).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU implementation of the same path
(the unmodified reference when /root/reference is mounted, else the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P, T, D = 196, 77, 512
L2_BYTES = 126 * 2 ** 20
METRIC = "sparc_infonce_fwd_bwd_pairs_per_s"
UNIT = "pairs/s"
# DRAM traffic of the dominant kernel per launch at config 2 (one `ncu --set full` capture, summary under profiles/)
TRAFFIC_BWD = 174.96e6
TRAFFIC_SRC = ("profiles/r2s_ncu_full_summary.csv (sparc_bwd3_kernel, B = 256, `ncu --set full` of tools/run_once.py 256): dram "
               "read 126.8 MB + write 48.2 MB per launch (algorithmic: inputs 71.6 + saved G 40.4 + logits 6.1 read, 71.6 written)")
SHARE_SRC = "profiles/r2s_launches.csv (sparc_bwd3 47 %, sparc_fwd3 30 %, global InfoNCE chain 22 % of the step's device time)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


def gpu_local_cpus(dev):
    """CPUs on the NUMA node the GPU hangs off (sysfs local_cpulist), or None.  Pinned staging buffers are first-touched by
    the allocating thread: allocating them from a CPU of the GPU's own node keeps the H2D copies of the end-to-end leg off the
    inter-socket link on multi-socket hosts.  (The boxes of this pool expose ONE node of 16 CPUs -- tools/numa_probe.py: 50-55
    GB/s either way -- so it is a no-op there; the e2e figure still varies from box to box, 114 k ... 194 k pairs/s.)"""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        return sorted(cpus) or None
    except Exception:
        return None


class numa_local:
    """with numa_local(dev): ... runs the block on the GPU's local CPUs (no-op when the topology is not exposed)."""

    def __init__(self, dev):
        self.cpus = gpu_local_cpus(dev)

    def __enter__(self):
        self.prev = os.sched_getaffinity(0)
        if self.cpus:
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        os.sched_setaffinity(0, self.prev)
        return False


def cfg(thr, s=1.0):
    return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0,
                                 inverse_temperature=s)


def flops_per_pair(Bg):
    """Algorithmic FLOPs per pair fwd+bwd: 12 T P D + 6 T^2 D + 6 Bg D."""
    return 12 * T * P * D + 6 * T * T * D + 6 * Bg * D


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def vit_l14_clip_shapes():
    """Parameter tensor shapes of HF CLIP ViT-L/14 (590 tensors, 427 616 513 elements)."""
    def tower(hidden, mlp, layers):
        s = []
        for _ in range(layers):
            s += [(hidden, hidden), (hidden,)] * 4            # k, v, q, out projections + biases
            s += [(hidden,), (hidden,)] * 2                   # layer_norm1/2 weight + bias
            s += [(mlp, hidden), (mlp,), (hidden, mlp), (hidden,)]
        return s
    text = [(49408, 768), (77, 768)] + tower(768, 3072, 12) + [(768,), (768,)]
    vision = [(1024,), (1024, 3, 14, 14), (257, 1024), (1024,), (1024,)] + tower(1024, 4096, 24) + [(1024,), (1024,)]
    shapes = [(1,)] + text + vision + [(768, 1024), (768, 768)]
    n = sum(int(torch.Size(s).numel()) for s in shapes)
    assert len(shapes) == 590 and n == 427616513, (len(shapes), n)
    return shapes


def bench_adamspd(dev, steps, warmup, pk):
    """AdamSPD full-model step on the ViT-L/14 CLIP tensor list (BASELINE config 5), fp32, random grads.
    `value` = algorithmic bytes / device time of cfa_adamspd_step (CUDA events around the launches on the launching
    stream); `ms_per_step_host_inclusive` = events around optimizer.step() starting from an idle GPU."""
    from clip_finegrained_alignment_b200 import AdamSPD, _lib
    shapes = vit_l14_clip_shapes()
    g = torch.Generator(device=dev).manual_seed(7)
    params = [torch.nn.Parameter(torch.randn(*s, device=dev, generator=g) * 0.02) for s in shapes]
    pre = [p.detach() + 1e-3 * torch.randn(*p.shape, device=dev, generator=g) for p in params]
    n_elts = sum(p.numel() for p in params)
    flat = torch.empty(n_elts + 4 * len(params), device=dev)      # grads are views of one buffer: one launch refills all
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += (p.numel() + 3) // 4 * 4                           # keep every view 16-byte aligned
    opt = AdamSPD([{"params": params, "pre": pre}], lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    numel = torch.tensor([p.numel() for p in params], dtype=torch.float64)

    def regrad():                                                  # fresh random grads so both SPD branches occur
        flat.normal_(0.0, 1e-3, generator=g)

    for _ in range(warmup):
        regrad(); opt.step()
    torch.cuda.synchronize(dev)
    host_ms, bytes_total = [], 0.0
    l0 = _lib.launch_count
    saved_events = _lib.kernel_events
    _lib.kernel_events = {"cfa_adamspd_step": []}
    for _ in range(steps):
        regrad()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); opt.step(); e1.record()
        torch.cuda.synchronize(dev)
        host_ms.append(e0.elapsed_time(e1))
        st = opt.last_step_stats.cpu().double()
        proj = (st[:, 0] > 0) & (st[:, 1] > 0)
        bytes_total += 32.0 * n_elts + 12.0 * float(numel[proj].sum())
    kev = _lib.kernel_events["cfa_adamspd_step"]
    _lib.kernel_events = saved_events
    launches = _lib.launch_count - l0
    ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)
    gbs = bytes_total / len(kev) / (ms * 1e-3) / 1e9
    # comparator: torch fused AdamW on the same tensors (28 B/elt)
    cmp_ms = None
    try:
        ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
        for q, p in zip(ref_params, params):
            q.grad = p.grad
        ref = torch.optim.AdamW(ref_params, lr=2e-5, weight_decay=0.1, fused=True)
        for _ in range(2):
            ref.step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ref.step()
        e1.record(); torch.cuda.synchronize(dev)
        cmp_ms = e0.elapsed_time(e1) / 3
        del ref, ref_params
    except Exception:
        pass
    # ---- AMP prologue folded into the step: amp_step vs unscale_ + clip_grad_norm_ + scaler.step
    amp = None
    try:
        scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
        scaler.scale(torch.zeros(1, device=dev))

        def timed(fn, n=3):
            ts = []
            for _ in range(n):
                regrad()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
            return min(ts)

        def fused():
            opt.amp_step(scaler, 1.0); scaler.update()

        def unfused():
            scaler.unscale_(opt); torch.nn.utils.clip_grad_norm_(params, 1.0); scaler.step(opt); scaler.update()

        fused(); unfused()
        f_ms, u_ms = timed(fused), timed(unfused)
        amp = {"amp_step_ms": round(f_ms, 4), "unscale_clip_step_ms": round(u_ms, 4),
               "amp_step_gbs": round((36.0 * n_elts + 12.0 * float(numel[proj].sum())) / (f_ms * 1e-3) / 1e9, 1),
               "note": "amp_step = cfa_adamspd_step_amp (one extra read of g: 36 B/elt + 12 B/elt projected, 4 launches, no "
                       "host sync); comparator = GradScaler.unscale_ + clip_grad_norm_ + GradScaler.step around the same "
                       "cfa_adamspd_step (finetuner.py:150-152)"}
    except Exception as e:      # keep the headline line even if the AMP leg fails
        amp = {"error": repr(e)[:200]}
    out = {"metric": "adamspd_step_hbm_gbs", "value": round(gbs, 1), "unit": "GB/s", "ms_per_step": round(ms, 4),
           "ms_per_step_host_inclusive": round(sum(host_ms) / len(host_ms), 4),
           "workload": "ViT-L/14 CLIP, 590 fp32 tensors, 427616513 params, random grads, lr 2e-5, wd 0.1",
           "algorithmic_bytes_per_step": bytes_total / len(kev), "gpu_launches_per_step": launches // max(1, steps),
           "roofline": {"bound": "hbm", "kernel": "adamspd_pass1 + adamspd_pass2", "achieved": round(gbs, 1), "peak": pk["hbm"],
                        "unit": "GB/s", "frac": round(gbs / pk["hbm"], 4),
                        "traffic": 16.0e9, "traffic_source": "profiles/r1d_ncu_adamspd.csv (dram read+write, both passes)",
                        "peak_source": pk["src"]},
           "torch_fused_adamw_ms": None if cmp_ms is None else round(cmp_ms, 4), "amp": amp}
    del opt, params, pre, flat
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
def _reference_modules():
    """The UNMODIFIED reference modules: oracle/_ref (made by oracle/make_ref.py, travels to the GPU box), refreshed from
    /root/reference when that is mounted (build container)."""
    from oracle import make_ref
    make_ref.make_ref(verbose=False)
    return make_ref.import_ref()


def make_cpu_runner(B):
    """The reference's CPU implementation of the path on a bounded sample: B pairs of the workload's shapes, fp32, all
    host threads, forward + autograd backward of the unmodified reference SPARCLoss (oracle/_ref); the oracle port only
    if oracle/_ref is absent."""
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(42)
    v = torch.randn(B, P, D, generator=g)
    l = torch.randn(B, T, D, generator=g)
    m = torch.ones(B, T, dtype=torch.bool)
    thr = 1.0 / P
    mods = _reference_modules()
    kind = "reference" if mods is not None else "port"
    if kind == "reference":
        mod = mods[0].SPARCLoss(cfg(thr))

        def run():
            vv = v.clone().requires_grad_(True)
            ll = l.clone().requires_grad_(True)
            mod(vv, ll, m)["total_loss"].backward()
    else:
        from oracle import losses_oracle as lo          # cpu_baseline leg: the checker, timed as the baseline

        def run():
            with torch.no_grad():
                lo.sparc_backward(lo.sparc_forward(v, l, m, thr, 1.0, 1.0, 1.0))
    what = "unmodified reference losses.py (autograd fwd + bwd)" if kind == "reference" else "oracle/losses_oracle.py port"
    sample = f"{B} pairs per step of the same shapes (P={P}, T={T}, D={D}), fp32, {what}"
    return run, kind, torch.get_num_threads(), sample


def gpu_eager(dev, B, steps=10):
    """Comparator named by BASELINE.md §4: the unmodified reference SPARCLoss in PyTorch eager mode ON THE B200 (bf16
    autocast as CUDA autocast would run it), fwd + autograd bwd at the bench batch."""
    mods = _reference_modules()
    if mods is None:
        return {"unavailable": "oracle/_ref not made"}
    try:
        crit = mods[0].SPARCLoss(cfg(1.0 / P)).to(dev)
        g = torch.Generator(device=dev).manual_seed(1)
        v = torch.randn(B, P, D, device=dev, generator=g).requires_grad_(True)
        l = torch.randn(B, T, D, device=dev, generator=g).requires_grad_(True)
        m = torch.ones(B, T, dtype=torch.bool, device=dev)

        def run():
            v.grad = None; l.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = crit(v, l, m)
            out["total_loss"].backward()

        for _ in range(3):
            run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(B / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4), "batch": B,
                "what": "unmodified reference SPARCLoss (oracle/_ref), torch eager + autocast(bf16) on the same B200, fp32 leaves"}
    except Exception as e:
        return {"error": repr(e)[:200]}


def cpu_baseline(B=32, budget_s=12.0):
    run, kind, cores, sample = make_cpu_runner(B)
    run()
    ts, t_start = [], time.perf_counter()
    while len(ts) < 3 or (time.perf_counter() - t_start < budget_s and len(ts) < 10):
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    return {"value": round(B / min(ts), 1), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": sample + f", best of {len(ts)}"}


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, rank 0 only."""
    if rank != 0:
        return
    B = 64
    run, kind, cores, sample = make_cpu_runner(B)
    for _ in range(args.warmup):
        run()
    ts = []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > 240:
            break
    sec = sum(ts) / len(ts)
    val = B / sec
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(ts), "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE config 2 shapes (ViT-B/16 SPARC + global InfoNCE fwd+bwd, P={P}, T={T}, "
                                   f"D={D}, thr=1/P, s=1), CPU, bounded sample of {B} pairs per step"},
            "cpu_baseline": {"value": round(val, 1), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="pairs per GPU; default: BASELINE config 2 (256) on one GPU, config 3 (1024 / GPU) with --gpus > 1")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32", "f16"])
    ap.add_argument("--patches", type=int, default=196, help="vision tokens P (BASELINE config 4: 576)")
    ap.add_argument("--dim", type=int, default=512, help="projection dim D (BASELINE config 4: 768)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: exchange of the gathered global InfoNCE (peer memory over NVLink, or NCCL all-gathers)")
    ap.add_argument("--cast-to-bf16", action="store_true", help="SPARCLoss(cast_to_bf16=True): fp16 / fp32 inputs on the tensor-core path")
    ap.add_argument("--no-graph", action="store_true", help="time the eager Python loop instead of CUDA-graph replays of the step")
    ap.add_argument("--no-adamspd", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global P, D
    P, D = args.patches, args.dim
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback)"
    args.warmup = max(args.warmup, 3)
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's debug output (its version banner is printed at every level) goes to stderr
        os.environ.setdefault("NCCL_DEBUG", os.environ.get("CFA_NCCL_DEBUG", "WARN"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from clip_finegrained_alignment_b200 import SPARCLoss, _lib
    pk = peaks()
    dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[args.dtype]
    B = args.batch if args.batch > 0 else (256 if world == 1 else 1024)
    Bg = B * world
    torch.manual_seed(42 + rank)
    mask_all = {}

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_run(B, steps, warmup, with_stage_events=True):
        """W untimed + K timed steps of SPARCLoss fwd + bwd at B pairs per GPU, inputs resident in HBM and rotated over
        enough sets to exceed 2 x L2; returns device-timed ms (max over ranks) and per-ABI-call device times."""
        in_bytes = B * (P + T) * D * torch.empty(0, dtype=dt).element_size()
        nbuf = max(2, -(-2 * L2_BYTES // in_bytes))            # rotate input sets: working set > 2 x L2
        vs = [torch.randn(B, P, D, device=dev).to(dt).requires_grad_(True) for _ in range(nbuf)]
        ls = [torch.randn(B, T, D, device=dev).to(dt).requires_grad_(True) for _ in range(nbuf)]
        mask = torch.ones(B, T, dtype=torch.bool, device=dev)
        mask_all[B] = mask
        crit = SPARCLoss(cfg(1.0 / P), gather=(True if args.collective == "peer" else "nccl") if world > 1 else False,
                         cast_to_bf16=args.cast_to_bf16)

        def step(i):
            v, l = vs[i % nbuf], ls[i % nbuf]
            v.grad = None; l.grad = None
            out = crit(v, l, mask)
            out["total_loss"].backward()
            return out["total_loss"]

        # initialisation, not warm-up: the first calls load the kernels (CUDA loads modules lazily, with a context
        # synchronisation each), size the allocator's pools and, with N > 1, create and map the peer exchange blocks
        init_steps = 10 if world > 1 else 2
        for i in range(init_steps):
            step(i)
        sync_all()
        for i in range(warmup):
            step(i)
        sync_all()
        # eager issue cost of one step (Python + autograd + 2 ctypes calls), measured without synchronising inside
        l0 = _lib.launch_count
        t_host = time.perf_counter()
        for i in range(min(steps, 10)):
            step(i)
        host_issue_ms = (time.perf_counter() - t_host) * 1e3 / min(steps, 10)
        sync_all()
        launches_per_step = (_lib.launch_count - l0) // min(steps, 10)
        # The timed region replays CUDA graphs of the same step (one graph per input set: forward, autograd backward and the
        # two library calls are captured through the public API with torch.cuda.graph, as a training loop would capture
        # its step): the device then runs back to back instead of waiting for ~0.2 ms of Python per 0.27 ms step.
        graphs, mode = [], "cuda_graph"
        if args.no_graph or world > 1:
            mode = "eager"
        else:
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for i in range(nbuf):
                        step(i)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                for i in range(nbuf):
                    vs[i].grad = None; ls[i].grad = None
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = crit(vs[i], ls[i], mask)
                        out["total_loss"].backward()
                    graphs.append((g, out["total_loss"]))
                for g, _ in graphs:
                    g.replay()
                torch.cuda.synchronize(dev)
                assert all(bool(torch.isfinite(t)) for _, t in graphs)
            except Exception as e:          # capture not possible: time the eager loop
                graphs, mode = [], "eager (graph capture failed: %s)" % repr(e)[:120]
                torch.cuda.synchronize(dev)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graphs:
            for i in range(steps):
                graphs[(warmup + i) % nbuf][0].replay()
        else:
            for i in range(steps):
                step(warmup + i)
        e1.record()
        sync_all()
        launches = launches_per_step * steps
        ms_total = e0.elapsed_time(e1)
        del graphs
        kev = {}
        if with_stage_events:
            # per-ABI-call device times: a separate short pass with CUDA events around every call (kept out of the timed
            # region: the extra event records cost host time)
            crit_stages = SPARCLoss(cfg(1.0 / P), gather=world > 1, fused_calls=False, cast_to_bf16=args.cast_to_bf16)      # same kernels, one call per stage
            for i in range(2):                           # untimed: kernels only this path launches (coefficient kernel) load lazily
                v, l = vs[i % nbuf], ls[i % nbuf]
                v.grad = None; l.grad = None
                crit_stages(v, l, mask)["total_loss"].backward()
            sync_all()
            _lib.kernel_events = {name: [] for name in _lib.LAUNCHES if name != "cfa_adamspd_step"}
            for i in range(min(10, steps)):
                v, l = vs[i % nbuf], ls[i % nbuf]
                v.grad = None; l.grad = None
                crit_stages(v, l, mask)["total_loss"].backward()
            sync_all()
            kev = _lib.kernel_events
            _lib.kernel_events = None
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        res = dict(B=B, ms_step=ms_total / steps, value=B * world * steps / (ms_total * 1e-3), launches=launches,
                   host_issue_ms=host_issue_ms, kev=kev, nbuf=nbuf, in_bytes=in_bytes, init_steps=init_steps, crit=crit, mode=mode)
        del vs, ls
        return res

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main_run = timed_run(B, args.steps, args.warmup)
    crit = main_run["crit"]
    mask = mask_all[B]
    kev, launches, host_issue_ms = main_run["kev"], main_run["launches"], main_run["host_issue_ms"]
    nbuf, in_bytes, init_steps = main_run["nbuf"], main_run["in_bytes"], main_run["init_steps"]
    ms_step, value = main_run["ms_step"], main_run["value"]
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in kev["cfa_sparc_bwd"])
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in kev["cfa_sparc_fwd"])

    # ---- end to end through the public API with HOST buffers (pinned): every step copies ITS inputs host->device and
    # reads its loss back device->host inside the timed region.  Like any input pipeline, the copy of step i+1 is
    # issued on a copy stream while step i computes (two device buffers); nothing is cached across steps.
    with numa_local(dev) as nl:
        hv = torch.randn(B, P, D).to(dt).pin_memory()
        hl = torch.randn(B, T, D).to(dt).pin_memory()
        hm = torch.ones(B, T, dtype=torch.bool).pin_memory()
    numa_note = f"; staging buffers first-touched on the GPU's NUMA node ({len(nl.cpus)} local CPUs)" if nl.cpus else ""
    bufs = [(torch.empty(B, P, D, dtype=dt, device=dev), torch.empty(B, T, D, dtype=dt, device=dev),
             torch.empty(B, T, dtype=torch.bool, device=dev)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def issue_copy(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])            # previous user of this buffer is done
            bufs[k][0].copy_(hv, non_blocking=True); bufs[k][1].copy_(hl, non_blocking=True)
            bufs[k][2].copy_(hm, non_blocking=True)
            copied[k].record(copy_stream)

    def e2e_run(n):
        losses_h = torch.empty(n, dtype=torch.float32).pin_memory()
        for k in range(2):
            consumed[k].record(main)
        issue_copy(0)
        for i in range(n):
            k = i & 1
            if i + 1 < n:
                issue_copy(k ^ 1)
            main.wait_event(copied[k])
            v = bufs[k][0].detach().requires_grad_(True); l = bufs[k][1].detach().requires_grad_(True)
            out = crit(v, l, bufs[k][2])
            out["total_loss"].backward()
            consumed[k].record(main)
            losses_h[i:i + 1].copy_(out["total_loss"].detach().reshape(1), non_blocking=True)   # D2H read of the result
        return losses_h

    e2e_run(3)
    sync_all()
    e2e_steps = max(5, min(args.steps, 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lh = e2e_run(e2e_steps)
    e1.record()
    sync_all()
    assert bool(torch.isfinite(lh).all())
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = Bg * e2e_steps / (float(t.item()) * 1e-3)
    clocks = sampler.stop() if sampler else None
    del bufs
    # N = 1 only: the same step at BASELINE config 3's per-GPU batch (B = 1024), so that the multi-GPU lines (which run
    # config 3, B = 1024 / GPU) have their own single-GPU reference in this line
    config3_n1 = None
    if world == 1 and args.batch == 0 and (P, D) == (196, 512):
        torch.cuda.empty_cache()
        r3 = timed_run(1024, max(5, args.steps // 3), 3, with_stage_events=False)
        config3_n1 = {"value": round(r3["value"], 1), "unit": UNIT, "ms_per_step": round(r3["ms_step"], 4), "batch_per_gpu": 1024,
                      "algorithmic_tflops": round(r3["value"] * flops_per_pair(1024) / 1e12, 2),
                      "note": "weak-scaling reference for the --gpus > 1 lines (BASELINE config 3 runs B = 1024 per GPU)"}
        del r3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    collective = "none (1 GPU)"
    if world > 1:
        from clip_finegrained_alignment_b200 import peer as _peer
        collective = ("peer memory: pooled embeddings and [lse | CE sums] packs read in place from the peers' HBM "
                      "(CUDA IPC over NVLink), 2 in-stream device barriers per step, no NCCL call on the step path"
                      if any(e.ok for e in _peer._EXCHANGES.values()) else "NCCL all-gather x2 per step")
    # roofline of the dominant kernel (cfa_sparc_bwd): algorithmic backward FLOPs of the fine-grained part per launch
    bwd_flops = B * (8 * T * P * D + 4 * T * T * D)
    ach = bwd_flops / (bwd_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"BASELINE config {(3 if (world > 1 and B == 1024) else 2) if (P, D) == (196, 512) else ('4' if (P, D) == (576, 768) else 'shapes')}: "
                               f"{'ViT-L/14@336' if P == 576 else 'ViT-B/16'} SPARC + global InfoNCE fwd+bwd, B={B}/GPU, P={P}, T={T}, "
                               f"D={D}, thr=1/P, s=1, all-True mask" + (", all-gathered global InfoNCE" if world > 1 else ""),
                   "global_batch": Bg, "collective": collective, "launch": main_run["mode"], "init_steps_before_warmup": init_steps, "l2": f"inputs rotated over {nbuf} sets ({nbuf * in_bytes >> 20} MiB > 2x L2)",
                   "algorithmic_flops_per_pair": flops_per_pair(Bg),
                   "algorithmic_tflops": round(value * flops_per_pair(Bg) / 1e12, 2),
                   "kernel_ms": {k: round(statistics.mean(a.elapsed_time(b) for a, b in ev), 4) for k, ev in kev.items() if ev}},
        "roofline": {"bound": "tensor", "kernel": "cfa_sparc_bwd", "achieved": round(ach, 2), "peak": pk["tf_sus"],
                     "unit": "TFLOP/s", "frac": round(ach / pk["tf_sus"], 5),
                     "traffic": TRAFFIC_BWD if (B == 256 and args.dtype == "bf16" and (P, D) == (196, 512)) else None,
                     "traffic_source": TRAFFIC_SRC, "share_of_device_time": SHARE_SRC,
                     "peak_source": pk["src"] + ", sustained bf16 GEMM", "algorithmic_flops_per_launch": bwd_flops},
        "e2e": {"value": round(e2e_val, 1), "unit": UNIT,
                "h2d_bytes_per_step": int(hv.numel() * hv.element_size() + hl.numel() * hl.element_size() + hm.numel()),
                "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "note": "pinned host buffers; H2D of step i+1 overlaps compute of step i (copy stream); PCIe-bound" + numa_note},
        "gpu_launches": launches, "host_issue_ms_per_step": round(host_issue_ms, 4), "clocks": clocks,
    }
    if config3_n1 is not None:
        line["config3_n1"] = config3_n1
    torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        line["gpu_eager"] = gpu_eager(dev, B)
        torch.cuda.empty_cache()
    if not args.no_adamspd:
        line["adamspd"] = bench_adamspd(dev, max(3, min(args.steps, 10)), 3, pk)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(32)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
