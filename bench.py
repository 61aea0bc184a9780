#!/usr/bin/env python
"""bench.py -- the measurement contract for the B200 KV-cache hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c4|c5|c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload (BASELINE.json configs[2], the configuration the metric "KV quant/dequant HBM GB/s at
1/2/4/8 B200" is quoted on; it fits one GPU):
    Llama-3-8B-shape synthetic KV cache -- 32 layers x {K, V}, each [B=16, H=8, T=32768, D=128] bf16
    (1 GiB) -- INT4 pack / unpack with a per-tensor scale, head-sharded over the N ranks.
One step = one pass over the whole batch: for every layer tensor  absmax (phase 1)  ->  [N > 1: ONE
allreduce(MAX) of the 64 absmax words over NCCL]  ->  quantize + nibble-pack (phase 2)  ->  unpack +
dequantize to bf16.  value = algorithmic bytes of the whole job / step time (max over ranks), inputs
resident in HBM (64 GiB >> 126 MB L2, so no flush is needed).  e2e = the same metric through the public
Python API (kvq_b200.quantize_int4_per_tensor_packed / dequantize_int4_per_tensor_packed) with HOST
buffers: pinned host -> device copy of every input and device -> host copy of the packed codes, the scale
and the dequantized tensor inside the timed region.

--impl reference times the reference's CPU path (the oracle port of its torch implementation, all host
cores) on a bounded sample of the same workload; rank 0 only.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "efficient-llm-inference_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_JSON_FD = 1


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


METRIC = "kv_quant_dequant_hbm_gbs"
UNIT = "GB/s"
C3 = dict(L=32, B=16, H=8, T=32768, D=128)


# ----------------------------------------------------------------------------- helpers


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram bytes (read+write) per launch of `kernel` from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ----------------------------------------------------------------------------- CPU reference arm


def cpu_reference_sample(cfg, seconds=12.0, T_slice=4096):
    """The reference's CPU path (oracle port, OpenMP over all host cores) on a T-slice of the same tensors:
    absmax -> INT4 quantize+pack -> unpack+dequantize to bf16, per tensor.  Returns (GB/s, cores, sample text, passes)."""
    import torch
    from oracle import kvq_oracle as o

    o.build()
    cores = os.cpu_count() or 1
    o.set_num_threads(cores)
    B, H, D = cfg["B"], cfg["H"], cfg["D"]
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, H, T_slice, D, generator=g, dtype=torch.float32).to(torch.bfloat16)
    n = x.numel()
    nbytes = 7 * n  # 2N absmax + 2.5N quant + 2.5N unpack
    o.quantize_int4_per_tensor_packed(x)  # warm (page faults, thread pool)
    passes, t0 = 0, time.perf_counter()
    while True:
        p, s, last = o.quantize_int4_per_tensor_packed(x)
        o.dequantize_int4_per_tensor_packed(p, s, last, torch.bfloat16)
        passes += 1
        el = time.perf_counter() - t0
        if el >= seconds or passes >= 4096:
            break
    gbs = nbytes * passes / el / 1e9
    sample = (f"{passes} passes over one [B={B},H={H},T={T_slice},D={D}] bf16 tensor (T sliced from {cfg['T']}); absmax + INT4 "
              f"quantize/pack + unpack/dequantize by the C oracle (OpenMP, {cores} threads), {el:.1f} s")
    return gbs, cores, sample, passes


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 3.0  # seconds of CPU work per "step" (a bounded sample of the workload)
    for _ in range(max(args.warmup, 0)):
        cpu_reference_sample(cfg, seconds=0.5)
    vals, t_all = [], time.perf_counter()
    sample = ""
    for _ in range(max(args.steps, 1)):
        gbs, cores, sample, _ = cpu_reference_sample(cfg, seconds=per_step)
        vals.append(gbs)
    value = statistics.mean(vals)
    ms = (time.perf_counter() - t_all) * 1e3 / max(args.steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(cfg, args.gpus, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "per step: " + sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(cfg, n_gpus, where):
    return {
        "workload": (f"Llama-3-8B-shape synthetic KV (L={cfg['L']} x {{K,V}}, B={cfg['B']}, H={cfg['H']} KV heads, T={cfg['T']}, "
                     f"D={cfg['D']}, bf16): INT4 pack/unpack, per-tensor scale, head-sharded over {n_gpus} GPU(s)"),
        "baseline_config": "BASELINE.json configs[2]",
        "step": ("absmax of all 64 layer tensors (1 launch) -> quantize+pack (1 launch) -> unpack+dequantize (1 launch)" if n_gpus == 1 or where != "gpu" else
                 "absmax (2 launches, 32 layer tensors each) -> allreduce-max of each half, overlapped with the other half's kernel -> "
                 "quantize+pack (2 launches) -> unpack+dequantize (1 launch)"),
        "algorithmic_bytes_per_element": 7.0,
        "l2_policy": "inputs (64 GiB whole job) are far larger than the 126 MB L2; no flush needed",
        "parallelism": f"kv-head shard x{n_gpus}" if where == "gpu" else "cpu",
    }


# ----------------------------------------------------------------------------- GPU arm (C3)


def run_c3(args, cfg):
    import torch
    import torch.distributed as dist

    import kvq_b200
    from kvq_b200 import _lib as L
    from kvq_b200.distributed import head_shard_range

    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    sm = ctypes.c_int()
    L.check(lib.kvq_device_info(ctypes.byref(sm), None, None), "kvq_device_info")

    Lyr, B, H, T, D = cfg["L"], cfg["B"], cfg["H"], cfg["T"], cfg["D"]
    lo, hi = head_shard_range(H, rank, world)
    Hl = hi - lo
    n_tensors = 2 * Lyr
    n_local = B * Hl * T * D
    n_global = B * H * T * D

    # synthetic inputs: one generator per (layer, K|V, head) so that 1-GPU and N-GPU data agree
    xs = []
    for layer in range(Lyr):
        for kv in range(2):
            x = torch.empty(B, Hl, T, D, device=dev, dtype=torch.bfloat16)
            for h in range(lo, hi):
                g = torch.Generator(device=dev).manual_seed(1234 + 1000 * layer + kv + 100000 * h)
                x[:, h - lo] = torch.randn(B, T, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            xs.append(x)
    packed = [torch.empty(B, Hl, T, D // 2, device=dev, dtype=torch.uint8) for _ in range(n_tensors)]
    outs = [torch.empty(B, Hl, T, D, device=dev, dtype=torch.bfloat16) for _ in range(n_tensors)]
    bits = torch.zeros(n_tensors, dtype=torch.int32, device=dev)
    scales = torch.empty(n_tensors, dtype=torch.bfloat16, device=dev)
    from kvq_b200.quantization import TensorTable
    tx, tp, to = TensorTable(xs), TensorTable(packed), TensorTable(outs)
    stream = L.stream_ptr(dev)
    dom_events = []  # (start, end) around every launch of the dominant kernel inside the timed region

    half = n_tensors // 2

    def step(timed: bool):
        # one launch per phase for the whole batch of 64 layer tensors (kvq_*_batched), one allreduce in between.  Head-sharded
        # (N > 1): the batch goes through absmax / quantize in two halves so that the allreduce of one half (NCCL's stream)
        # overlaps the other half's kernel -- SURVEY 8e: "overlap by launching phase-1 for tensor i+1 before phase-2 of tensor i"
        n_launch = 3
        if world > 1 and not os.environ.get("KVQ_BENCH_NO_ALLREDUCE"):  # (env: debugging aid, not a valid bench line)
            works = []
            for lo_t, cnt in ((0, half), (half, n_tensors - half)):
                L.check(lib.kvq_absmax_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, bits.data_ptr() + 4 * lo_t, stream),
                        "kvq_absmax_batched")
                works.append(dist.all_reduce(bits[lo_t: lo_t + cnt], op=dist.ReduceOp.MAX, async_op=True))  # the only collective on the path
            for (lo_t, cnt), wk in zip(((0, half), (half, n_tensors - half)), works):
                wk.wait()  # stream-level wait
                L.check(lib.kvq_quant_batched(tx.ptrs.data_ptr() + 8 * lo_t, L.BF16, cnt, n_local, 4 | L.ABSMAX_COVERS_INPUT,
                                              bits.data_ptr() + 4 * lo_t, 1e-8, tp.ptrs.data_ptr() + 8 * lo_t,
                                              scales.data_ptr() + 2 * lo_t, None, stream), "kvq_quant_batched")
            n_launch = 5
        else:
            L.check(lib.kvq_absmax_batched(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, bits.data_ptr(), stream), "kvq_absmax_batched")
            L.check(lib.kvq_quant_batched(tx.ptrs.data_ptr(), L.BF16, n_tensors, n_local, 4 | L.ABSMAX_COVERS_INPUT, bits.data_ptr(), 1e-8, tp.ptrs.data_ptr(),
                                          scales.data_ptr(), None, stream), "kvq_quant_batched")
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        L.check(lib.kvq_dequant_batched(tp.ptrs.data_ptr(), 4, n_tensors, n_local, scales.data_ptr(), L.BF16, to.ptrs.data_ptr(),
                                        L.BF16, stream), "kvq_dequant_batched")
        if timed:
            e1.record()
            dom_events.append((e0, e1))
        return n_launch

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    sync_all()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        launches += step(True)
    ev1.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t)
    ms_step = ms_total / args.steps
    job_bytes = 7.0 * n_global * n_tensors
    value = job_bytes / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (dequant_flat_kernel<bf16, INT4>): algorithmic N*(0.5+2) bytes per launch
    dom_ms = statistics.mean(a.elapsed_time(b) for a, b in dom_events)
    peak, peak_src = measured_peak()
    dom_bytes = 2.5 * n_local * n_tensors  # one launch covers all layer tensors
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9

    # parity spot check inside the run: codes of tensor 0 against the exact formula on a sample
    x0 = xs[0]
    idx = torch.randint(0, x0.numel(), (1 << 16,), device=dev)
    s0 = torch.clamp(bits[:1].view(torch.float32).cpu() / 7.0, min=1e-8).to(dev)
    want = torch.clamp(torch.round(x0.view(-1)[idx].float() / s0), -8, 7).to(torch.int64)
    byte = packed[0].view(-1)[idx // 2].to(torch.int64)
    got = torch.where(idx % 2 == 0, byte >> 4, byte & 15) - 8
    assert torch.equal(got, want), "bench parity spot check failed"

    del xs, packed, outs, x0, x, tx, tp, to  # the e2e leg brings its data from the host
    torch.cuda.empty_cache()
    decode = run_decode(args, dev, rank, world) if not args.skip_decode else None
    e2e = run_e2e(args, cfg, dev, Hl, world) if not args.skip_e2e else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0,
                                                                          "d2h_bytes_per_step": 0, "_ms": 1.0, "_job_bytes": 0.0}
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        gbs, cores, sample, _ = cpu_reference_sample(cfg, seconds=12.0)
        cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if world > 1 and not args.skip_e2e:
        # aggregate e2e over ranks: bytes of all ranks / max time
        t = torch.tensor([e2e["_ms"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["value"] = e2e["_job_bytes"] * world / (float(t) * 1e-3) / 1e9
        e2e["h2d_bytes_per_step"] *= world
        e2e["d2h_bytes_per_step"] *= world
    e2e.pop("_ms"), e2e.pop("_job_bytes")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(cfg, world, "gpu"),
            "pct_of_8TBs_per_gpu": 100.0 * value / world / 8000.0,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "dequant_flat_kernel<bf16,int4> (unpack+dequantize)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms, "launches_timed": len(dom_events),
                         "traffic": (lambda t: None if t is None else t * n_tensors * n_local / (16 * 8 * 32768 * 128))(ncu_traffic("dequant_flat_int4_bf16"))},
            "cpu_baseline": cpu,
            "decode": decode,
            "sm_count": sm.value,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


C5 = dict(L=80, B=8, H=8, T=65536, D=128)  # Llama-3-70B-shape KV (BASELINE.json configs[4]), INT8-K / INT4-V


def run_decode(args, dev, rank, world):
    """Second half of the metric ("decode tok/s"): fused dequant decode attention over the C5 cache, head-sharded (no
    collective: heads are independent).  One step = the attention of all 80 layers for a batch of 8 sequences at 64k
    context; tok/s = B / step time (max over ranks).  Captured in a CUDA graph so the GPU time is not hidden behind
    Python; reported for the MHA shape (one query head per KV head: bandwidth-bound) and for Llama-3-70B's real GQA
    (8 query heads per KV head: FMA-bound on CUDA cores)."""
    import torch
    import torch.distributed as dist

    import kvq_b200
    from kvq_b200.distributed import head_shard_range

    c = C5
    lo, hi = head_shard_range(c["H"], rank, world)
    Hl, B, T, D, Lyr = hi - lo, c["B"], c["T"], c["D"], c["L"]
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    kcs = [torch.randint(-127, 128, (B, Hl, T, D), device=dev, dtype=torch.int8, generator=g) for _ in range(Lyr)]
    vcs = [torch.randint(0, 256, (B, Hl, T, D // 2), device=dev, dtype=torch.uint8, generator=g) for _ in range(Lyr)]
    ksc = torch.rand(T, device=dev, generator=g) * 0.01 + 0.01
    vsc = torch.rand(T, device=dev, generator=g) * 0.01 + 0.01
    ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    out = {}
    nbytes = Lyr * (B * Hl * T * (D + D // 2) + 8 * T)
    for name, G in (("mha", 1), ("gqa8", 8)):
        q = torch.randn(B, Hl * G, 1, D, device=dev, dtype=torch.bfloat16, generator=g)

        def step():
            for l in range(Lyr):
                kvq_b200.decode_attention_quantized(q, kcs[l], ksc, 8, vcs[l], vsc, 4, workspace=ws)

        step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        out[name] = {"tok_s": B / (ms * 1e-3), "ms_per_step": ms, "gbs_whole_job": nbytes * world / (ms * 1e-3) / 1e9,
                     "q_heads_per_kv_head": G}
        del graph
    out["config"] = (f"Llama-3-70B-shape synthetic KV (L={Lyr}, {c['H']} KV heads, D={D}) batch {B} x {T} context, INT8-K / INT4-V "
                     f"per-token-scaled cache, fused dequant decode attention, head-sharded over {world} GPU(s); BASELINE.json configs[4]")
    out["launches_per_step"] = 2 * Lyr
    del kcs, vcs
    torch.cuda.empty_cache()
    return out


def run_e2e(args, cfg, dev, Hl, world):
    """Same metric through the public API with host buffers.  A ring of pinned host tensors feeds H2D copies on a
    copy stream, the kernels run through kvq_b200.quantize_int4_per_tensor_packed / dequantize_int4_per_tensor_packed on a
    compute stream, results (packed codes, scale, dequantized tensor) go back to pinned host memory on a third stream;
    every one of the 64 layer tensors of a step makes that round trip inside the timed region."""
    import torch

    import kvq_b200

    Lyr, B, T, D = cfg["L"], cfg["B"], cfg["T"], cfg["D"]
    n_tensors = 2 * Lyr
    shape = (B, Hl, T, D)
    n = B * Hl * T * D
    ring = 2
    g = torch.Generator().manual_seed(99)
    host_in = [torch.empty(shape, dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    for h in host_in:
        h.view(torch.int16).random_(-16000, 16000, generator=g)  # finite bf16 bit patterns
    host_packed = [torch.empty((B, Hl, T, D // 2), dtype=torch.uint8).pin_memory() for _ in range(ring)]
    host_out = [torch.empty(shape, dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    host_scale = [torch.empty((), dtype=torch.bfloat16).pin_memory() for _ in range(ring)]
    dev_in = [torch.empty(shape, dtype=torch.bfloat16, device=dev) for _ in range(ring)]
    s_h2d, s_cmp, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(ring)]
    ev_done = [torch.cuda.Event() for _ in range(ring)]
    ev_free = [torch.cuda.Event() for _ in range(ring)]
    ev_out_free = [torch.cuda.Event() for _ in range(ring)]

    def one_step():
        keep = []
        for i in range(n_tensors):
            r = i % ring
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_free[r])          # compute finished reading dev_in[r]
                dev_in[r].copy_(host_in[r], non_blocking=True)
                ev_in[r].record(s_h2d)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[r])
                packed, scale, last = kvq_b200.quantize_int4_per_tensor_packed(dev_in[r])
                out = kvq_b200.dequantize_int4_per_tensor_packed(packed, scale, last, torch.bfloat16)
                ev_free[r].record(s_cmp)
                ev_done[r].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_done[r])
                host_packed[r].copy_(packed, non_blocking=True)
                host_scale[r].copy_(scale, non_blocking=True)
                host_out[r].copy_(out, non_blocking=True)
                packed.record_stream(s_d2h), out.record_stream(s_d2h), scale.record_stream(s_d2h)
            keep.append((packed, out))
            if len(keep) > ring:
                keep.pop(0)
        return 4 * n_tensors

    steps = max(1, min(args.steps, 2))
    for e in ev_free:
        e.record(s_cmp)
    one_step()  # warm (allocator, pinned pages)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(torch.cuda.current_stream(dev))
    s_h2d.wait_event(t0), s_cmp.wait_event(t0), s_d2h.wait_event(t0)
    wall0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    cur = torch.cuda.current_stream(dev)
    for s in (s_h2d, s_cmp, s_d2h):
        e = torch.cuda.Event()
        e.record(s)
        cur.wait_event(e)
    t1.record(cur)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    wall_ms = (time.perf_counter() - wall0) * 1e3 / steps
    job_bytes = 7.0 * n * n_tensors
    h2d = n_tensors * n * 2
    d2h = n_tensors * (n // 2 + 2 + n * 2)
    return {"value": job_bytes / (ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": ms, "wall_ms_per_step": wall_ms, "steps": steps,
            "api": "kvq_b200.quantize_int4_per_tensor_packed + dequantize_int4_per_tensor_packed on pinned host buffers "
                   f"(ring of {ring}), 3 streams", "_ms": ms, "_job_bytes": job_bytes}


def main():
    # stdout carries exactly ONE line (the JSON): libraries that print to fd 1 (NCCL's version banner does) go to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kvq", choices=["kvq", "reference"])
    ap.add_argument("--layers", type=int, default=None, help="override the number of layers (debug only; not a valid bench line)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: leave out the CPU baseline leg")
    ap.add_argument("--skip-decode", action="store_true", help="profiling runs only: leave out the decode-attention (C5) leg")
    args = ap.parse_args()
    cfg = dict(C3)
    if args.layers:
        cfg["L"] = args.layers
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    run_c3(args, cfg)


if __name__ == "__main__":
    main()
