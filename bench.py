#!/usr/bin/env python
"""Benchmark of the fp16 C2C FFT hot path (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): batched 1-D C2C fp16 FFT, N = 16384,
batch 4096 per GPU, iid N(0,1) planar input in the reference batch layout [RE_b | IM_b]
(src/base/DataHandler.h:105-114).  One "step" = one transform of the whole batch.
  value     : whole-job GFLOP/s (5*N*log2(N) per transform), inputs resident in HBM, CUDA-event
              timed on the launching stream, max over ranks.  Weak scaling: every GPU
              transforms its own 4096-transform shard, no collective on the data path.
  e2e       : same metric through the C ABI with HOST buffers (tfft_exec_host: pinned H2D +
              kernel + D2H every step), the reference's CopyDataHostToDevice -> ComputeFFT ->
              CopyResultsDeviceToHost sequence.
  roofline  : algorithmic bytes (8*N*batch per launch) / average launch duration
              against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline : the fp64 host FFT oracle (oracle/tfft_oracle.cpp) on a bounded sample, all cores.
--impl reference runs the UNMODIFIED reference kernels (oracle/_ref/libtfft_ref.so, built from
/root/reference by oracle/Makefile) through the reference's own API on a bounded sample of the
same workload; if that library or a GPU is missing it times the host fp64 oracle instead.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))

N = 16384
BATCH = int(os.environ.get("TFFT_BENCH_BATCH", "4096"))   # developer knob; the contract workload is 4096
METRIC = "fp16 C2C FFT GFLOP/s (5*N*log2N), batched 1-D N=16384 x 4096 per GPU"
FLOP_PER_TRANSFORM = 5.0 * N * 14


def read_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic():
    """dram bytes per launch from the committed ncu summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """SM clock / throttle-reason sampler (NVML, ~1 kHz) running during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.smax, self.thread, self.err = None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ent = vis.split(",")[self.index].strip()
                if ent.isdigit():
                    idx = int(ent)
                else:
                    uuid = ent
            self.h = nv.nvmlDeviceGetHandleByUUID(uuid) if uuid else nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv = nv
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:  # noqa
                self.err = repr(e)
                break
            time.sleep(0.0005)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
               "samples": len(sm), "source": "NVML polled during the timed region"}
        if self.err:
            out["error"] = self.err
        return out


def cpu_baseline(sample_transforms=512, min_seconds=6.0):
    """fp64 host FFT (oracle port) on a bounded sample of the workload, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    re, im = O.gauss_fixture(N, sample_transforms, seed=1234)
    re, im = re.astype(np.float64), im.astype(np.float64)
    O.fft_f64(re[:8], im[:8])
    reps, t0 = 0, time.perf_counter()
    while True:
        O.fft_f64(re, im)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= 4000:
            break
    gflops = FLOP_PER_TRANSFORM * sample_transforms * reps / dt / 1e9
    return {"value": round(gflops, 3), "unit": "GFLOP/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{sample_transforms} of {BATCH} transforms x {reps} reps, fp64 radix-2 host FFT "
                      f"(oracle/tfft_oracle.cpp), {dt:.1f} s"}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args):
    import numpy as np
    import torch
    import tfft
    rank, world, local = dist_setup(args.gpus)
    K, W = args.steps, max(args.warmup, 3)
    # synthetic shard of this rank, generated on the device (counter-based seed per rank)
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + rank)
    d_in = torch.randn(BATCH * 2 * N, generator=g, device="cuda", dtype=torch.float32).to(torch.float16)
    d_out = torch.empty_like(d_in)
    plan = tfft.NativePlan(N, BATCH)
    passes = plan.info["passes"]

    def step():
        plan.exec(d_in, d_in[N:], d_out, d_out[N:], 2 * N, 2 * N)

    for _ in range(W):
        step()
    sampler = ClockSampler(local)
    barrier(world)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world)
    clocks = sampler.stop()
    ms_step = ms_total / K

    # end to end through the C ABI with pinned host buffers
    h_in = torch.empty(BATCH * 2 * N, dtype=torch.float16).pin_memory()
    h_in.copy_(d_in)
    h_out = torch.empty(BATCH * 2 * N, dtype=torch.float16).pin_memory()
    e2e_steps = max(3, min(K, 10))
    for _ in range(2):
        plan.exec_host(h_in.numpy(), h_out.numpy())
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.exec_host(h_in.numpy(), h_out.numpy())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier(world)
    e2e_ms_step = max_over_ranks(e2e_s * 1e3 / e2e_steps, world)

    # light self-check of the timed result (device vs e2e path must agree bit for bit)
    same = bool(torch.equal(d_out[: 64 * 2 * N].cpu(), h_out[: 64 * 2 * N]))

    if rank != 0:
        return
    flop_step = FLOP_PER_TRANSFORM * BATCH * world
    alg_bytes = 8.0 * N * BATCH * passes
    peak, peak_src = read_peak()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": round(flop_step / (ms_step * 1e-3) / 1e9, 1), "unit": "GFLOP/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_step, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 storage, f32 accumulate", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: batched 1-D C2C fp16 FFT N=16384 x batch 4096 per GPU, "
                               "planar [RE_b|IM_b] layout, 1/N scaled",
                   "n": N, "batch_per_gpu": BATCH, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "working set 512 MiB per step > 126 MB L2 (inputs larger than L2)"},
        "e2e": {"value": round(flop_step / (e2e_ms_step * 1e-3) / 1e9, 1), "unit": "GFLOP/s",
                "h2d_bytes_per_step": 4 * N * BATCH, "d2h_bytes_per_step": 4 * N * BATCH,
                "ms_per_step": round(e2e_ms_step, 3), "api": "tfft_exec_host (C ABI, pinned host buffers; chunked H2D / transform / D2H pipeline on three streams)"},
        "gpu_launches": K * passes,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": read_traffic(), "peak_source": peak_src,
                     "kernel": "tfft::fft_unit_kernel_2slot<4,5,5>", "algorithmic_bytes_per_launch": int(alg_bytes),
                     "hbm_gbs_p1": round(8.0 * N * BATCH / (ms_step * 1e-3) / 1e9, 1)},
        "clocks": clocks, "self_check_device_eq_e2e": same,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)


def run_reference(args):
    """The reference's own implementation (unmodified kernels + API) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    K, W = args.steps, max(args.warmup, 3)
    sample = 256   # transforms per step: the batch overload creates (and leaks) one stream per transform
    line = {"impl": "reference", "metric": METRIC, "unit": "GFLOP/s", "n_gpus": 1, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: batched 1-D C2C fp16 FFT N=16384 x batch 4096 per GPU, "
                                   "planar [RE_b|IM_b] layout, 1/N scaled",
                       "n": N, "batch_per_gpu": BATCH,
                       "sample": f"{sample} of {BATCH} transforms per step (reference batch overload creates one "
                                 "stream per transform and never destroys it, src/base/ComputeFFT.h:167-173)"}}
    use_gpu = False
    try:
        import torch
        use_gpu = torch.cuda.is_available() and O.ref_lib() is not None
    except Exception:
        use_gpu = False
    cpu = cpu_baseline()
    if use_gpu:
        K = min(K, 20)
        k_ms, e_ms = O.ref_bench_gpu(N, sample, K, min(W, 5), mode=0, use_batch_api=True)
        flop = FLOP_PER_TRANSFORM * sample
        kv = flop / (float(np.mean(k_ms)) * 1e-3) / 1e9
        ev = flop / (float(np.mean(e_ms)) * 1e-3) / 1e9
        line.update({"value": round(kv, 2), "ms_per_step": round(float(np.mean(k_ms)), 4), "dtype": "f16",
                     "steps": K,
                     "reference_kind": "unmodified reference CUDA kernels (oracle/_ref/libtfft_ref.so built from "
                                       "/root/reference/src/base), default plan Mode_256, its batch ComputeFFT",
                     "e2e": {"value": round(ev, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * N * sample,
                             "d2h_bytes_per_step": 4 * N * sample,
                             "api": "DataBatchHandler::CopyDataHostToDevice + ComputeFFT + CopyResultsDeviceToHost"},
                     "cpu_baseline": cpu, "gpu_launches": 0})
    else:
        line.update({"value": cpu["value"], "dtype": "f64", "ms_per_step": None,
                     "reference_kind": "host fp64 FFT oracle (the reference has no CPU implementation and "
                                       "oracle/_ref or a GPU is unavailable)",
                     "cpu_baseline": cpu,
                     "e2e": {"value": cpu["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
