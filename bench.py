#!/usr/bin/env python
"""Benchmark of the Sub-GC hot path (BASELINE.json metric: captions/sec, 36-node sub-graphs, 20-token decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode greedy|topk|beam]

One step = one pass of the whole path (fusion -> GCN -> sGPN -> NMS -> prepare -> 20-token decode) over one batch of
128 synthetic images per GPU (BASELINE config 2: 128 images x 36 nodes x 2048-d, one full 36-node sub-graph kept per
image, greedy decode).  N > 1: launched by torchrun, every rank decodes its own 128 images (weak scaling, no
data-path collective; NCCL only carries the barrier and the max-over-ranks time).

Prints ONE JSON line on rank 0:
  value        whole-job captions/s with inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e          same metric through the public model API from pinned HOST buffers, H2D + D2H inside the timed region
  roofline     decode loop (subgc_decode_sample): algorithmic bytes per launch / CUDA-event duration vs measured HBM peak
  cpu_baseline the oracle port of the reference algorithm (torch CPU ops) on this box's host cores, same workload
`--impl reference` times that CPU path alone (the reference itself is pure PyTorch; oracle/ restates it and is pinned to
it by tests/golden — the one place besides tests/smoke where bench.py executes oracle/).
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
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))

import torch  # noqa: E402

from subgc import synth  # noqa: E402
from subgc.config import Dims, make_opt  # noqa: E402

IMAGES_PER_GPU = 128
SEED = 2019  # test.py:21 of the reference
METRIC = "captions/sec (36-node sub-graphs, 20-token decode)"

# algorithmic work of the decode loop, fp32 storage (SURVEY §8d / BASELINE.md §4)
W_BYTES = 4 * (4000 * 3000 + 4000 * 1000 + 4000 * 2000 + 4000 * 1000 + 512 * 1000 + 9488 * 1000) \
    + 4 * (8000 + 8000 + 512 + 9488)                       # LSTM + h2att + logit weights and biases: 152.1 MB
ROW_BYTES = 4 * (36 * 1000 + 36 * 512 + 8 * 1000 + 1000 + 1000)  # att + p_att + state r/w + fc + embed row: 257.7 KB
ROW_FLOPS = 76.13e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0))), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                try:
                    rs = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                flags = ["Active" if rs & names[n] else "Not Active" for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                self.rows.append([str(sm), str(mx), "0"] + flags)
            except Exception:
                pass
            time.sleep(0.002)

    def _run(self):
        try:
            return self._run_nvml()
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(d, rank):
    data = synth.make_test_inputs(d, SEED + rank, n_images=IMAGES_PER_GPU, per_half=1, ragged=False, ragged_edges=False)
    return data


def cpu_reference(d, sd, data, mode, steps, warmup, threads):
    """The reference algorithm on host cores (oracle port), whole workload per step.  Returns captions/s, ms/step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import subgc_oracle as O
    torch.set_num_threads(threads)
    kw = dict(use_nms=True, iou_thres=0.75, max_subgraphs=1)
    if mode == "topk":
        kw.update(topk=True, temp=0.6, k=3)
    elif mode == "beam":
        kw.update(beam_size=5)
    n = None
    with torch.no_grad():
        for _ in range(warmup):
            n = O.sample(sd, d, data, **kw)["seq"].shape[0]
        t0 = time.perf_counter()
        for _ in range(steps):
            n = O.sample(sd, d, data, **kw)["seq"].shape[0]
        dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps * 1e3, n


def reference_arm(d, sd, data, mode, steps, warmup, threads, n_sample, check_seq=None):
    """The UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py) on host cores, driven as it can be driven: one image
    per call (models/lib/gpn.py:84).  Each step decodes the first `n_sample` images of the workload.  Returns None when baseline/_ref is
    absent (then the oracle port is the CPU arm)."""
    sys.path.insert(0, ROOT)
    from baseline import ref_runner as R
    if not R.available():
        return None
    torch.set_num_threads(threads)
    names = list(synth.SAMPLE_ARG_ORDER)
    model = R.load(d, sd, make_opt, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if mode == "topk" else 0)
    opt = {"beam_size": 5 if mode == "beam" else 1}
    images = range(min(n_sample, data["att_feats"].shape[0]))
    for _ in range(warmup):
        R.run(model, data, names, images[:2], opt)
    n_tot, t_tot, seqs = 0, 0.0, None
    for _ in range(steps):
        n, t, seqs = R.run(model, data, names, images, opt)
        n_tot += n; t_tot += t
    out = {"value": n_tot / t_tot, "ms_per_step": t_tot / steps * 1e3, "captions_per_step": n_tot // steps}
    if check_seq is not None and mode == "greedy":
        ref_seq = torch.cat([x.cpu() for x in seqs])
        out["tokens_equal_gpu"] = bool(torch.equal(ref_seq, check_seq[:ref_seq.shape[0]].cpu()))
    return out


def bench_train(args, d, dev, rank, world, warmup):
    """BASELINE config 5 (secondary line): Sub_GC_Kar training step = LossWrapper forward + backward (dropout on) + gradient
    all-reduce (NCCL) on `--train-images` images per GPU (5 sentences each, 17 teacher-forced steps).  No optimiser step."""
    import torch.distributed as dist
    from subgc import parallel
    from subgc.model import LossWrapper, setup
    sd = synth.make_state_dict(d, SEED)
    model = setup(make_opt(d))
    model.load_state_dict(sd)
    model.to(dev).train()
    lw = LossWrapper(model, None)
    data = synth.make_train_inputs(d, SEED + rank, n_images=args.train_images, gpn_batch=2, ragged=True, ragged_edges=True)
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    call = (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
            data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
    params = list(model.parameters())
    from subgc import _lib
    from subgc.optim import ClipAdam
    L = _lib.lib()
    reducer = parallel.GradReducer(world) if world > 1 else None
    model.grad_reducer = reducer          # per-bucket all-reduce started from inside the hand-written backward (overlapped)

    def step():
        for p in params:
            p.grad = None
        out = lw(*call)
        (out["lang_loss"] + out["gpn_loss"]).backward()
        return out

    for _ in range(warmup):
        out = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    c0 = L.subgc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exposed = []
    e0.record()
    for _ in range(args.steps):
        out = step()
        if reducer is not None:
            exposed.append(reducer._exposed)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = (L.subgc_launch_count() - c0) / args.steps
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    # the optimiser step that follows in train.py:163-164 (not part of BASELINE config 5's metric): fused global-norm clip + Adam
    opt = ClipAdam([p for p in params], 5e-4, clip_norm=10.0)
    for _ in range(2):
        opt.step()
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(5):
        opt.step()
    o1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0]) / args.steps
        sent = args.train_images * 5 * world
        exp_ms = sum(a.elapsed_time(b) for a, b in exposed) / max(len(exposed), 1) if exposed else 0.0
        n_grad = sum(p.grad.numel() for p in params if p.grad is not None)
        print(json.dumps({"metric": "training sentences/sec (Sub_GC_Kar step: forward + LanguageModelCriterion + backward + grad all-reduce)",
                          "value": sent / (ms * 1e-3), "unit": "sentences/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"BASELINE config 5: {args.train_images} images/GPU ({args.train_images * 5} sentences), 17 "
                                                 "teacher-forced steps, dropout 0.5, grads all-reduced (fp32), no optimiser step in the timed region"},
                          "subgc_launches_per_step": launches,
                          "allreduce": {"bytes": 4 * n_grad, "buckets": 3, "overlapped": world > 1,
                                        "exposed_ms_per_step": exp_ms, "exposed_frac": exp_ms / ms if ms else None,
                                        "note": "GPU time the compute stream waits for the bucket collectives after the last backward kernel "
                                                "(+ the 1/world scaling); buckets: decoder, prepare, sGPN+GCN+fusion, each started when final"},
                          "optimizer_step_ms": o0.elapsed_time(o1) / 5,
                          "lang_loss": float(out["lang_loss"]), "gpn_loss": float(out["gpn_loss"])}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="greedy", choices=["greedy", "topk", "beam", "train"])
    ap.add_argument("--train-images", type=int, default=32, help="images per GPU for --mode train (BASELINE config 5: 256 / 8 GPUs)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    d = Dims()
    cfg_no = {"greedy": "2", "beam": "3 (beam_size 5)", "topk": "4 shard (top-k sampling, k=3, temp 0.6)", "train": "5"}[args.mode]
    config = {"workload": f"BASELINE config {cfg_no}: {IMAGES_PER_GPU} images/GPU x 36 nodes x 2048-d, 64 edges, 1 full sub-graph kept per image "
                          f"(NMS 0.75/max 1), {args.mode} 20-token decode, V=9487",
              "images_per_gpu": IMAGES_PER_GPU, "rows_per_gpu": IMAGES_PER_GPU, "decode": args.mode, "weights": "random init (synthetic), fp32 (+ split-fp16 packed copies of the same 4 bytes per weight for the tensor cores)",
              "l2": "per-step working set (280 MB weights + 83 MB inputs + activations) exceeds the 126 MB L2; no explicit flush"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        sd = synth.make_state_dict(d, SEED)
        data = make_inputs(d, 0)
        n_sample = 8 if args.mode == "beam" else 32
        ref = reference_arm(d, sd, data, args.mode, max(args.steps, 1), args.warmup, cores, n_sample)
        if ref is not None:
            val, ms = ref["value"], ref["ms_per_step"]
            pval, pms, pn = cpu_reference(d, sd, data, args.mode, 1, 1, cores)
            cb = {"value": val, "unit": "captions/s", "cores": cores, "kind": "reference",
                  "sample": f"{ref['captions_per_step']} images of the workload per step, one call per image as the reference requires "
                            f"(models/lib/gpn.py:84); unmodified reference from baseline/_ref, torch {torch.__version__} CPU, {cores} threads",
                  "port": {"value": pval, "ms_per_step": pms, "sample": f"oracle port, all {pn} images in one batched call (an extension the "
                                                                        "reference does not have), 1 step"}}
        else:
            val, ms, n = cpu_reference(d, sd, data, args.mode, max(args.steps, 1), args.warmup, cores)
            cb = {"value": val, "unit": "captions/s", "cores": cores, "kind": "port",
                  "sample": f"whole workload: {n} captions per step (oracle port of the reference's PyTorch path, "
                            f"torch {torch.__version__} CPU, {cores} threads); baseline/_ref is not installed"}
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "captions/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": val, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from subgc import _lib
    from subgc.model import setup
    if args.mode == "train":
        return bench_train(args, d, dev, rank, world, warmup)

    sd = synth.make_state_dict(d, SEED)
    model = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if args.mode == "topk" else 0))
    model.load_state_dict(sd)
    model.to(dev).eval()
    data = make_inputs(d, rank)
    names = [k for k in synth.SAMPLE_ARG_ORDER]
    from subgc import compact
    # loader-shaped tuple with the tensors the kernels never read left out (nothing is uploaded in vain), and the compact wire format
    lean = dict(zip(names, compact.needed_only(*[data[k] for k in names])))
    host = {k: (lean[k].pin_memory() if lean[k] is not None else None) for k in names}
    resident = {k: (host[k].to(dev) if host[k] is not None else None) for k in names}
    host_cb = compact.compact_batch(*[data[k] for k in names]).packed(pin=True)   # one pinned buffer: the upload is ONE host->device copy
    opt = {"beam_size": 5 if args.mode == "beam" else 1}
    if args.mode == "topk":
        opt["seed"] = SEED
    loader_bytes = sum(data[k].numel() * data[k].element_size() for k in names if data[k] is not None)
    lean_bytes = sum(t.numel() * t.element_size() for t in host.values() if t is not None)
    h2d_bytes = host_cb.nbytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return model(*[resident[k] for k in names], opt=opt, mode="sample")

    # end to end: every step's inputs start in pinned HOST memory and its results end in host memory.  The H2D copy of step
    # i+1 is issued on a copy stream while step i computes (two device-side input sets); the timed region contains all of it.
    # Two input formats are timed: "compact" (subgc.compact wire format, mode='sample_compact': the headline e2e) and "lean" (the
    # reference-signature call, unread tensors not uploaded).
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [{k: (torch.empty_like(resident[k]) if resident[k] is not None else None) for k in names} for _ in range(2)]
    dev_cb = [host_cb.packed(device=dev, copy=False) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i, fmt):
        with torch.cuda.stream(copy_stream):
            if fmt == "compact":
                dev_cb[i % 2].copy_(host_cb, non_blocking=True)
            else:
                for k in names:
                    if host[k] is not None:
                        dev_in[i % 2][k].copy_(host[k], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def step_e2e(i, last, fmt):
        if not last:
            prefetch(i + 1, fmt)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        if fmt == "compact":
            out = model(dev_cb[i % 2], opt=opt, mode="sample_compact")
        else:
            out = model(*[dev_in[i % 2][k] for k in names], opt=opt, mode="sample")
        hr = getattr(model, "last_host_results", None)
        if hr is not None:
            # the call itself read its whole result pack back (one device-to-host copy, see TopDownModel._sample_dyn): the host tensors
            res = [hr["seq"], hr["seqLogprobs"], hr["subgraph_score"], hr["keep_ind"]]
        else:
            res = [t.to("cpu", non_blocking=True) if t.is_cuda else t for t in out]
        torch.cuda.synchronize()
        return res

    def time_e2e(fmt):
        n_warm = 2 * max(warmup, 3)   # each of the two device-side input sets goes through eager -> graph capture -> replay first
        prefetch(0, fmt)
        for i in range(n_warm):
            r = step_e2e(i, i == n_warm - 1, fmt)
        barrier()
        t0 = time.perf_counter()
        prefetch(n_warm, fmt)                                   # all K host->device copies happen inside the timed region
        for i in range(n_warm, n_warm + args.steps):
            r = step_e2e(i, i == n_warm - 1 + args.steps, fmt)
        ms = (time.perf_counter() - t0) * 1e3                   # host wall clock: every step ends with a device synchronisation
        barrier()
        return ms, r

    L = _lib.lib()
    with torch.no_grad():
        for _ in range(warmup):
            out = step_resident()
        n_rows = out[0].shape[0]
        assert n_rows == IMAGES_PER_GPU, n_rows
        # ---- timed region 1: device-resident inputs -----------------------------------------------------------------
        sampler = ClockSampler(local_rank) if rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        from subgc.model import _DecodePlan
        launches0 = L.subgc_launch_count() + _DecodePlan.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = step_resident()
        e1.record()
        barrier()
        launches = L.subgc_launch_count() + _DecodePlan.replayed_launches - launches0   # direct launches + kernels inside graph replays
        ms_total = e0.elapsed_time(e1)
        # ---- stage split and the decode launch time behind `roofline`: a second pass over the same inputs with CUDA events between
        # the stages (the timed pass above replays the whole step as one graph, which has no place for events); same kernels
        model.stage_events = []
        for _ in range(3):
            step_resident()
        torch.cuda.synchronize()
        model.stage_events = []
        import ctypes as _C
        kernel_ms = []   # duration of the persistent decode kernel alone: CUDA events recorded around its launch on the launching stream
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        if args.mode != "beam":   # third pass, launches issued eagerly (a captured launch cannot carry events): the kernel alone
            graphs_on, model.use_graphs = model.use_graphs, False
            events_kept, model.stage_events = model.stage_events, []
            L.subgc_mega_timing(1, None)
            for i in range(args.steps + 2):
                step_resident()
                t_k = _C.c_float(-1.0)
                L.subgc_mega_timing(1, _C.byref(t_k))
                if t_k.value > 0 and i >= 2:
                    kernel_ms.append(t_k.value)
            L.subgc_mega_timing(0, None)
            model.use_graphs, model.stage_events = graphs_on, events_kept
        torch.cuda.synchronize()
        stage_ms = {}
        for name, a, b in model.stage_events:
            stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b)
        model.stage_events = None
        # ---- timed region 2: end to end from pinned host memory -----------------------------------------------------
        e2e_ms_total, res = time_e2e("compact")
        lean_ms_total, res_lean = time_e2e("lean")
        assert all(torch.equal(a, b) for a, b in zip(res, res_lean)), "compact and loader-shaped inputs must give identical results"
        d2h_bytes = sum(t.numel() * t.element_size() for t in res)
        clocks = sampler.stop() if sampler else None

    times = torch.tensor([ms_total, e2e_ms_total, lean_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, lean_ms_total = float(times[0]), float(times[1]), float(times[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    steps_exec = int(model.last_steps.item()) if model.last_steps is not None else d.seq_length + 1
    total_caps = n_rows * world * args.steps
    value = total_caps / (ms_total * 1e-3)
    e2e_value = total_caps / (e2e_ms_total * 1e-3)
    hbm_peak, tf_peak, peak_kind = load_peaks()
    line = {"metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms_total / args.steps,
                    "input": "compact wire format of subgc.compact (class ids, node lists + lengths, one copy per image) in pinned host "
                             "memory, model(batch, mode='sample_compact')",
                    "loader_shaped": {"value": total_caps / (lean_ms_total * 1e-3), "unit": "captions/s", "h2d_bytes_per_step": lean_bytes,
                                      "ms_per_step": lean_ms_total / args.steps,
                                      "input": "reference-signature call, tensors the kernels never read not uploaded "
                                               f"(the loaders' full tuple is {loader_bytes} bytes)"}},
            "gpu_launches": int(launches), "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "stage_note": "stage times and roofline.launch_ms come from a second pass with CUDA events between the stages; `value` is the "
                          "whole step replayed as one CUDA graph (encoder .. persistent decode kernel, no host round trip before the results)",
            "decode_steps_executed": steps_exec}
    if "decode" in stage_ms:
        t_dec = stage_ms["decode"] / args.steps * 1e-3           # the decode stage: fc_pre contraction + the decode loop
        if kernel_ms and args.mode != "beam":                      # greedy / top-k: the loop is ONE kernel, timed alone (stage minus fc_pre etc.)
            t_dec = sum(kernel_ms) / len(kernel_ms) * 1e-3
        algo_steps = d.seq_length                                 # 20 algorithmic steps per caption
        dec_rows = n_rows * (5 if args.mode == "beam" else 1)     # decoder rows in flight (beam search: 5 beams per sub-graph)
        algo_bytes = algo_steps * (W_BYTES + dec_rows * ROW_BYTES)
        algo_flops = algo_steps * dec_rows * ROW_FLOPS
        hbm = {"achieved": algo_bytes / t_dec / 1e9, "peak": hbm_peak, "unit": "GB/s"}
        # the contractions run as 3 tcgen05 kind::f16 products per fp32 product (split-fp16 operands): the tensor-pipe roofline counts them
        tens = {"achieved": 3 * algo_flops / t_dec / 1e12, "peak": tf_peak, "unit": "TFLOP/s"}
        traffic, traffic_note = None, "no ncu capture found under profiles/"
        tp = os.path.join(ROOT, "profiles", "r02_decode_traffic.json")
        if os.path.isfile(tp) and args.mode in ("greedy", "topk"):
            tj = json.load(open(tp))
            traffic, traffic_note = tj["dram_bytes_per_decode_loop"], tj["note"]
        bound = "tensor" if tens["achieved"] / tens["peak"] > hbm["achieved"] / hbm["peak"] else "hbm"
        main = tens if bound == "tensor" else hbm
        kernel = ("mega_decode_kernel: the decode loop of subgc_decode_sample as one persistent cooperative launch (20 x [att-LSTM, cell, h2att, "
                  "attention, lang-LSTM, cell, logit, select])") if args.mode != "beam" else \
                 ("decode loop of subgc_decode_beam: 20 x [att-LSTM, cell, h2att, attention, lang-LSTM, cell, logit contraction (h3_gemm_kernel), "
                  "per-row top-b, beam step], 640 rows, one CUDA graph")
        line["roofline"] = {"bound": bound, "kernel": kernel, "achieved": main["achieved"], "peak": main["peak"], "unit": main["unit"],
                            "frac": main["achieved"] / main["peak"], "traffic": traffic, "traffic_note": traffic_note,
                            "peak_source": peak_kind + " (MEASURED_PEAKS.json: hbm_gbs / bf16_tflops_sustained)",
                            "algorithmic_bytes_per_launch": algo_bytes, "algorithmic_flops_per_launch": algo_flops, "launch_ms": t_dec * 1e3,
                            "launch_ms_note": ("mean of CUDA events recorded around each mega_decode_kernel launch (subgc_mega_timing), "
                                               f"{len(kernel_ms)} launches" if kernel_ms and args.mode != "beam" else "decode stage of the per-stage pass"),
                            "hbm": dict(hbm, frac=hbm["achieved"] / hbm["peak"]),
                            "tensor": dict(tens, frac=tens["achieved"] / tens["peak"],
                                           note="3 x algorithmic flops: every fp32 product is three fp16 tensor-core products (hi.hi, hi.lo, lo.hi) "
                                                "with fp32 accumulation in TMEM")}
    if "encode" in stage_ms and "roofline" in line:
        # front-end: tensor-bound.  Executed contraction work per image with folded units and the node <- edges layer
        # aggregated before its contraction: fusion 37 x (2048 + 300) x 1024, L0 37 x 1024 x 2048, L1 2 x 37 x 1024 x 1024,
        # sGPN / prepare as BASELINE.md counts them; every fp32 product = 3 fp16 tensor-core products
        enc_flops = IMAGES_PER_GPU * 2.0 * (37 * 2348 * 1024 + 37 * 1024 * 2048 + 2 * 37 * 1024 * 1024)
        prep_flops = n_rows * 132.2e6
        t_enc, t_prep = stage_ms["encode"] / args.steps * 1e-3, stage_ms.get("prepare", 0.0) / args.steps * 1e-3
        line["roofline"]["front_end"] = {
            "bound": "tensor", "peak": tf_peak, "unit": "TFLOP/s",
            "encode": {"ms": t_enc * 1e3, "executed_flops": enc_flops, "achieved": 3 * enc_flops / t_enc / 1e12, "frac": 3 * enc_flops / t_enc / 1e12 / tf_peak},
            "prepare": {"ms": t_prep * 1e3, "executed_flops": prep_flops, "achieved": 3 * prep_flops / max(t_prep, 1e-9) / 1e12,
                        "frac": 3 * prep_flops / max(t_prep, 1e-9) / 1e12 / tf_peak},
            "note": "stage time includes the non-contraction kernels (splits, segment means, pooling); per-kernel tensor-pipe activity: "
                    "profiles/r02d_ncu_encoder.md"}
    if not args.no_cpu_baseline:
        n_sample = 16 if args.mode == "beam" else IMAGES_PER_GPU
        ref = reference_arm(d, sd, data, args.mode, 1, 1, cores, n_sample, check_seq=out[0])
        cval, cms, cn = cpu_reference(d, sd, data, args.mode, args.cpu_steps if ref is None else 1, 1, cores)
        port = {"value": cval, "unit": "captions/s", "ms_per_step": cms,
                "sample": f"oracle port of the reference's PyTorch path, all {cn} images in one batched call, {cores} host threads"}
        if ref is not None:
            line["cpu_baseline"] = {"value": ref["value"], "unit": "captions/s", "cores": cores, "kind": "reference", "ms_per_step": ref["ms_per_step"],
                                    "sample": f"{ref['captions_per_step']} images of this workload, one call per image (models/lib/gpn.py:84), "
                                              f"unmodified reference from baseline/_ref on {cores} host threads, 1 pass after a 2-image warm-up",
                                    "tokens_equal_gpu": ref.get("tokens_equal_gpu"), "port": port}
        else:
            line["cpu_baseline"] = dict(port, cores=cores, kind="port")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
