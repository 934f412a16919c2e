#!/usr/bin/env python
"""bench.py — headline benchmark: aggregate IQ Msamples/s of the stereo FM receiver hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own CPU code on the host cores)

Workload = BASELINE.json configs[1]: mode 0 stereo FM (pilot BPF + PLL + 38 kHz mix), a batch of 256
independent synthetic 8-bit IQ streams per GPU (weak scaling: N GPUs carry N x 256 streams, partitioned by
stream index, no data-path collective).  One "step" = one pass of the whole receiver over the batch:
256 streams x 48 blocks (1.024 s of signal each, 1.26 GB of input per GPU — larger than the 126 MB L2).  The synthetic
streams are periodic over that length (tones, pilot and FM phase complete whole cycles), so feeding the buffer again at every
step continues every stream without a jump: the receivers carry their state from step to step as they would on live input.

One JSON line on stdout (rank 0):
  value     device-resident throughput: inputs already in HBM, CUDA events on the launch stream, max over ranks
  e2e       same metric through Pipeline.process_host(): pinned HOST input -> H2D -> kernels -> D2H PCM, per step
  roofline  the fused front-end kernel (largest share of the arithmetic) against the FP32 roofline, timed live
            with CUDA events around each of its launches; `kernels` lists every kernel's time share so the
            PLL's share (latency-bound serial recurrence) is visible next to it
  cpu_baseline  the reference's own code (oracle/_ref, else the C port) on all host cores, bounded sample
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE, STEREO = 0, 1
RDS = False                      # --rds: BASELINE configs[3], the stereo receiver plus the RDS path
STREAMS_PER_GPU = 256
BLOCKS_PER_STREAM = 48           # 48 x 21.33 ms = 1.024 s per stream: every component of the synthetic signal completes whole cycles in it
METRIC = "aggregate_iq_msamples_per_s_stereo_fm"
UNIT = "Msamples/s"
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.45: SMs x lanes x 2 flop x max SM clock
# Measured ceiling of the BIT-EXACT multiply-add (FFMA2(x,h,-0) then FADD2: two packed instructions on the fmaheavy pipe per
# pair of MACs), tools/ubench.cu "exact2" on this pool's B200 (profiles/ubench_r1b.txt): 36.3 TFLOP/s = 0.49 of the FMA peak.
EXACT_MAC_CEILING_TFLOPS = 36.3
# DRAM traffic of k_frontend_stream from the round-2 `ncu --set full` capture of the whole-job launch, profiles/r2_ncu_k_frontend_stream.md:
# (1323.959 + 257.409) MB for 256 streams x 48 blocks x 51200 pairs  ->  bytes per IQ pair (algorithmic: 2.4).  DRAM counters cannot
# be read without a profiler attached, so this figure is carried from that capture and scaled to the launch; per-kernel traffic of a
# whole step is in profiles/r2_step_traffic.md.
NCU_FRONTEND_DRAM_BYTES_PER_PAIR = (1323.959e6 + 257.408512e6) / (256 * 48 * 51200)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(kind, mode, stereo, seeds, n_blocks, barrier, q):
    import numpy as np
    import oracle
    import importlib.util
    spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "3dy4-real-time-software-defined-radio-_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    lib = oracle.load(kind)
    m = lib.mode_params(mode)
    streams = [synth.make_stream(mode, n_blocks * m.block_size // 2, s) for s in seeds]
    lib.pipeline(mode, stereo, streams[0][:m.block_size], want=("pcm",))      # warm the code path
    barrier.wait()
    t0 = time.time()
    n = 0
    for iq in streams:
        lib.pipeline(mode, stereo, iq, want=("pcm",))
        n += iq.size // 2
    q.put((t0, time.time(), n))


def cpu_throughput(n_streams, n_blocks, cores):
    """One reference pipeline per stream, `cores` worker processes in flight (the reference replay swaps
    std::cin's buffer, so workers are processes, not threads).  Returns (Msamples/s, kind, pairs, seconds)."""
    import oracle
    kind = "ref" if oracle.have_ref() else "oracle"
    ctx = mp.get_context("fork")
    barrier, q = ctx.Barrier(cores), ctx.Queue()
    seeds = [[65 + s for s in range(w, n_streams, cores)] for w in range(cores)]
    procs = [ctx.Process(target=_cpu_worker, args=(kind, MODE, STEREO, seeds[w], n_blocks, barrier, q)) for w in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    t0, t1, pairs = min(r[0] for r in res), max(r[1] for r in res), sum(r[2] for r in res)
    return pairs / (t1 - t0) / 1e6, ("reference" if kind == "ref" else "port"), pairs, t1 - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_streams, n_blocks = 4 * cores, 24          # per step: 4 streams per core x 0.51 s of signal (~0.8 s of wall per step)
    for _ in range(args.warmup):
        cpu_throughput(cores, 2, cores)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        v, kind, pairs, dt = cpu_throughput(n_streams, n_blocks, cores)
        vals.append(v)
        secs += dt
    value = sum(vals) / len(vals)
    sample = "%d streams x %d blocks (%.2f s of signal each) per step, one reference pipeline per stream, %d worker processes" % (
        n_streams, n_blocks, n_blocks * 0.021333, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * secs / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "note": "each worker runs the reference's frontend() then backend() per block on ONE thread (oracle/ref_replay.cpp), not main()'s two "
                                 "short-lived threads per block: that favours the reference slightly (no thread spawn/join per 21 ms block)"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (NVML, ~100 Hz)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ready, self.armed = threading.Event(), False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            self.ready.set()                         # NVML is up: the timed region may start
            while not self.stop_flag:
                if self.armed:                       # only samples taken while the timed region runs count
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:      # NVML missing: report that, do not fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)
            self.ready.set()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def workload_config(overlap=None):        # overlap: how the GPU arm issues its calls (None: the reference arm, where it does not apply)
    return {
        "workload": "%smode %d %s FM (8-bit IQ -> IF -> %s int16 PCM), %d independent synthetic streams per GPU x %d blocks"
                    % ("BASELINE configs[1]: " if (MODE, STREAMS_PER_GPU, STEREO) == (0, 256, 1) else "", MODE, "stereo" if STEREO else "mono",
                       "L/R" if STEREO else "mono", STREAMS_PER_GPU, BLOCKS_PER_STREAM),
        "mode": MODE, "stereo": bool(STEREO), "rds": RDS, "streams_per_gpu": STREAMS_PER_GPU, "blocks_per_stream": BLOCKS_PER_STREAM,
        "input_bytes_per_gpu": STREAMS_PER_GPU * BLOCKS_PER_STREAM * {0: 102400, 1: 81920, 2: 160000, 3: 128000}[MODE],
        "l2": "inputs are larger than the 126 MB L2 (see input_bytes_per_gpu); no flush needed",
        "arithmetic": "front end, pilot/stereo BPF and PLL bit-exact to the reference (unfused f32; the PLL's f64 libm results reproduced exactly); "
                      "audio resamplers fused f32 (PCM within 1 LSB)",
        "parallelism": "streams partitioned across GPUs, one process per GPU, no collective on the data path",
        "calls": None if overlap is None else "one dy4_pipeline_process call per step; " + (
            "steps are queued back to back and OVERLAP on the device (DY4_FLAG_PIPELINED: step k+1's FIR kernels run beside step k's PLL loops), "
            "one dy4_pipeline_flush before the closing event - every step's PCM is complete inside the timed region" if overlap else
            "every step is joined to the stream before the next starts"),
    }


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import numpy as np
    import torch
    import dy4_b200
    from dy4_b200 import shard

    rank, local_rank, world = shard.init_process_group()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    m = dy4_b200.mode_params(MODE)
    S, nb = STREAMS_PER_GPU, BLOCKS_PER_STREAM
    lo, hi = shard.stream_range(S * world, world, rank)
    assert hi - lo == S
    pairs_per_step = S * nb * m.block_size // 2
    n_if, n_audio = nb * m.if_per_block, nb * m.audio_per_block

    # synthetic input, generated on the GPU; stream s of the whole job uses seed 65+s
    periodic = dy4_b200.synth.whole_cycles(MODE, nb * m.block_size // 2)       # 48 blocks in mode 0: yes
    d_iq = dy4_b200.synth.make_batch_torch(MODE, min(S, 512), nb * m.block_size // 2, base_seed=65 + lo, device=dev, rds=RDS, periodic=periodic)
    if S > 512:                                      # very large batches: tile 512 distinct streams (generation time, not a kernel matter)
        d_iq = d_iq.repeat((S + 511) // 512, 1)[:S].contiguous()
    overlap = bool(STEREO) and not args.no_overlap  # consecutive steps overlap on the device (DY4_FLAG_PIPELINED); one flush before the closing event
    pipe = dy4_b200.Pipeline(MODE, STEREO, S, device=local_rank, rds=RDS, pipelined=overlap)
    nch = 2 if STEREO else 1
    out = {"pcm": torch.empty((S, n_audio * nch), dtype=torch.int16, device=dev)}
    max_steps_between_resets = max(1, int(55.0 / (nb * (m.block_size / 2) / m.rf_Fs)))   # float PLL sample counter saturates at 2^24 (~69.9 s)

    def step(i):
        if i % max_steps_between_resets == 0:
            pipe.reset()
        pipe.process(d_iq, n_blocks=nb, want=("pcm",), out=out)
        if RDS:
            pipe.rds_drain(raw=True)                 # symbols / bits / frame-sync events of this step, to the host

    for i in range(args.warmup):                     # the streams start here (step 0 resets) ...
        step(i)
    pipe.flush()
    torch.cuda.synchronize(dev)

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    pipe.profile(True)
    pipe.profile_get(reset=True)
    launches0 = dy4_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.ready.wait(timeout=10)
    shard.barrier()
    torch.cuda.synchronize(dev)
    sampler.armed = True
    e0.record()
    for i in range(args.steps):                      # ... and simply continue through the timed region
        step(args.warmup + i)
    pipe.flush()                                     # pipelined calls: the stream waits for every step's last kernel
    e1.record()
    torch.cuda.synchronize(dev)
    sampler.armed = False
    shard.barrier()
    ms = e0.elapsed_time(e1)
    launches = dy4_b200.launch_count() - launches0
    prof = pipe.profile_get(reset=True)
    pipe.profile(False)
    loop_sms, rest_sms = pipe.sm_partition()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # one step alone on an idle device, queue to flush: the LATENCY of a call (the timed region above measures throughput)
    lat = []
    for i in range(2):
        torch.cuda.synchronize(dev)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        step(args.warmup + args.steps + i)
        pipe.flush()
        l1.record()
        torch.cuda.synchronize(dev)
        lat.append(l0.elapsed_time(l1))
    ms_max = shard.max_over_ranks(ms, dev)
    total_launches = int(shard.sum_over_ranks(launches, dev))
    value = world * pairs_per_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public host API: pinned host input, H2D + kernels + D2H PCM every step -----
    h_iq = torch.empty((S, nb * m.block_size), dtype=torch.uint8).pin_memory()
    h_iq.copy_(d_iq)
    # Output: ONE int16 array [all streams of the job][samples] in shared pinned host memory; each rank's device->host copies land
    # in its slice, so after the barrier rank 0 holds the whole batch's PCM: that IS the path's host-side gather.
    shared = shard.SharedRows("dy4_bench_pcm_%s" % os.environ.get("MASTER_PORT", str(os.getpid())), S * world, n_audio * nch, np.int16, rank, world)
    h_pcm = torch.from_numpy(shared.mine)
    e2e_steps = max(1, min(args.steps, 8))
    pipe.reset()
    pipe.process_host(h_iq, n_blocks=nb, want=("pcm",), out={"pcm": h_pcm}, chunk_blocks=args.chunk_blocks)   # warm-up (allocates staging)
    pipe.sync()
    torch.cuda.synchronize(dev)
    shard.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):                       # the streams continue from step to step, as in the device-resident leg
        pipe.process_host(h_iq, n_blocks=nb, want=("pcm",), out={"pcm": h_pcm}, chunk_blocks=args.chunk_blocks)
    pipe.sync()                                      # overlapped calls: every upload, kernel and download of the steps above is done
    torch.cuda.synchronize(dev)
    shard.barrier()                                  # every rank's slice has landed: rank 0 now holds the gathered PCM
    e2e_ms = shard.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    gathered_ok = bool(rank != 0 or all(np.any(shared.all[r * S]) for r in range(world)))      # a row of every rank is there
    e2e_value = world * pairs_per_step * e2e_steps / (e2e_ms * 1e-3) / 1e6
    # the two paths from the same starting state: identical PCM?
    pipe.reset()
    pipe.process_host(h_iq, n_blocks=nb, want=("pcm",), out={"pcm": h_pcm}, chunk_blocks=args.chunk_blocks)
    pipe.reset()
    pipe.process(d_iq, n_blocks=nb, want=("pcm",), out=out)
    pipe.flush()
    torch.cuda.synchronize(dev)
    pcm_matches = bool(torch.equal(h_pcm.to(dev), out["pcm"]))
    gather_info = {"kind": "shared pinned host array (/dev/shm, cudaHostRegister): each rank's D2H copies land in its slice of rank 0's buffer",
                   "pinned": shared.pinned, "bytes_total": int(S * world * n_audio * nch * 2), "inside_e2e_timing": True, "all_slices_present": gathered_ok}
    shared.close()

    # ---- the timed path against the oracle (rank 0, two streams x 4 blocks from a fresh start): same bytes in, PCM out -----
    oracle_check = None
    if rank == 0 and not args.no_cpu and not RDS:
        import oracle
        chk = oracle.load("ref") if oracle.have_ref() else oracle.load("oracle")
        nbc = min(4, nb)
        small = d_iq[:2, :nbc * m.block_size].contiguous()
        pc = dy4_b200.Pipeline(MODE, STEREO, 2, device=local_rank, pipelined=overlap)        # as the timed path: two calls, overlapped
        h1 = nbc // 2
        parts = [pc.process(small[:, :h1 * m.block_size], n_blocks=h1, want=("pcm",))["pcm"],
                 pc.process(small[:, h1 * m.block_size:], n_blocks=nbc - h1, want=("pcm",))["pcm"]] if h1 else [pc.process(small, n_blocks=nbc, want=("pcm",))["pcm"]]
        pc.flush()
        torch.cuda.synchronize(dev)
        got = torch.cat(parts, 1).cpu().numpy()
        pc.close()
        worst = 0
        for s_ in range(2):
            ref = chk.pipeline(MODE, STEREO, small[s_].cpu().numpy(), want=("pcm",))["pcm"]
            worst = max(worst, int(np.abs(got[s_].astype(np.int64) - ref.astype(np.int64)).max()))
        oracle_check = {"checker": chk.kind, "streams": 2, "blocks": nbc, "pcm_max_abs_diff_lsb": worst, "ok": worst <= 1}

    # ---- the same kernels in whole-job launches, not overlapped with the PLL (one sub-chunk per step) ------------
    # In the timed region above the job is cut into sub-chunks of 1, 2, 4, ... blocks whose FIR kernels run beside the
    # PLL of their predecessor: good for the step, but small concurrent launches understate what a kernel can do.
    iso = None
    if rank == 0 and not RDS:
        pipe_iso = dy4_b200.Pipeline(MODE, STEREO, S, device=local_rank, debug_rows=True)
        for i in range(2):
            pipe_iso.process(d_iq, n_blocks=nb, want=("pcm",), out=out)
        torch.cuda.synchronize(dev)
        pipe_iso.profile(True)
        pipe_iso.profile_get(reset=True)
        pipe_iso.reset()
        for i in range(2):
            pipe_iso.process(d_iq, n_blocks=nb, want=("pcm",), out=out)
        torch.cuda.synchronize(dev)
        iso = pipe_iso.profile_get(reset=True)
        pipe_iso.close()

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- per-kernel breakdown and the roofline of the front-end kernel ---------------------------------------
    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"
    # algorithmic bytes / MACs per IQ pair, SURVEY.md §8(d) and DESIGN.md §5
    rd = float(m.rf_decim)                           # IQ pairs per IF sample
    ad = rd * m.audio_decim / m.audio_upsample       # IQ pairs per audio sample
    PLL_TABLE = bool(STEREO) and S <= int(os.environ.get("DY4_PLL_TABLE_MAX", "1024" if RDS else "4096"))     # dy4_pipeline.cu: ensure_workspace
    alg = {
        "frontend": {"bytes": 2.0 + 4 / rd, "mac": 2 * 101 / rd},
        "twin_bpf": {"bytes": 4 / rd + 8 / rd, "mac": 2 * 101 / rd},
        # direct loop: pilot 4 + reciprocal 8 in, phase row 8 out per IF sample; table-driven loop: 32-byte table row in, phaseEst 4 out
        "pll": {"bytes": (36 if PLL_TABLE else 20) / rd, "mac": 0.0},
        "audio": {"bytes": (12 if STEREO else 4) / rd + 4 / ad, "mac": (2 if STEREO else 1) * 101 / ad},
        "tails": {"bytes": 0.0, "mac": 0.0},
        # direct: reciprocals (4 in, 8 out) and NCO row (8 in, 4 out); table-driven: prediction (4 in, 4 out), table (8 in, 32 out), NCO (4 in, 4 out)
        "pll_aux": {"bytes": (56 if PLL_TABLE else 24) / rd, "mac": 0.0},
        # RDS path: two 101-tap band-pass filters; PLL rows; 19/120 resampler + RRC on I and Q
        "rds_bpf": {"bytes": 16 / rd, "mac": 2 * 101 / rd},
        "rds_pll": {"bytes": 28 / rd, "mac": 0.0},
        "rds_baseband": {"bytes": 12 / rd, "mac": 2 * 2 * 101 * 19 / 120 / rd},
    }
    total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {}
    for k, v in prof.items():
        if not v["launches"]:
            continue
        avg = v["ms"] / v["launches"]
        pairs_per_launch = pairs_per_step * args.steps / v["launches"]
        kernels[k] = {
            "launches": v["launches"], "avg_ms": round(avg, 4), "share": round(v["ms"] / total_kernel_ms, 4),
            "GBps": round(alg[k]["bytes"] * pairs_per_launch / (avg * 1e-3) / 1e9, 1),
            "fp32_TFLOPs": round(2 * alg[k]["mac"] * pairs_per_launch / (avg * 1e-3) / 1e12, 2),
        }
    fe = kernels["frontend"]
    isolated = {}
    for k, v in (iso or {}).items():
        if v["launches"] and alg[k]["mac"]:
            avg = v["ms"] / v["launches"]
            tf = 2 * alg[k]["mac"] * pairs_per_step / (avg * 1e-3) / 1e12
            isolated[k] = {"avg_ms": round(avg, 4), "fp32_TFLOPs": round(tf, 2), "frac_of_fma_peak": round(tf / FP32_PEAK_TFLOPS_NOMINAL, 4)}
            if k in ("frontend", "twin_bpf") and STEREO:        # the two bit-exact kernels
                isolated[k]["frac_of_exact_ceiling"] = round(tf / EXACT_MAC_CEILING_TFLOPS, 4)
    roofline = {
        "kernel": "k_frontend_stream (packed uint8 IQ -> 101-tap decimating FIR on I,Q in transposed form -> FM discriminator)",
        "bound": "fp32", "achieved": fe["fp32_TFLOPs"], "peak": round(FP32_PEAK_TFLOPS_NOMINAL, 2), "unit": "TFLOP/s",
        "frac": round(fe["fp32_TFLOPs"] / FP32_PEAK_TFLOPS_NOMINAL, 4),
        "peak_source": "148 SMs x 128 FP32 lanes x 2 x 1.965 GHz; tools/ubench measures 71.6 TFLOP/s FFMA on this pool",
        "note": ("bit-exact arithmetic (unfused multiply then add) costs two packed FP32 instructions per pair of MACs, both on the fmaheavy pipe: "
                 "the measured ceiling of that instruction pair is %.1f TFLOP/s (tools/ubench.cu exact2, profiles/ubench_r1b.txt), frac %.2f" % (
                     EXACT_MAC_CEILING_TFLOPS, EXACT_MAC_CEILING_TFLOPS / FP32_PEAK_TFLOPS_NOMINAL)) if STEREO else
                "mono receiver: nothing chaotic downstream, so the front end's multiply-add is fused (one FFMA2 per tap and I/Q pair); the ceiling is the FMA peak",
        "exact_mac_ceiling": EXACT_MAC_CEILING_TFLOPS if STEREO else None,
        "frac_of_exact_ceiling": round(fe["fp32_TFLOPs"] / EXACT_MAC_CEILING_TFLOPS, 4) if STEREO else None,
        "hbm": {"achieved": fe["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": round(fe["GBps"] / hbm_peak, 4), "peak_source": hbm_src},
        "share_of_step": fe["share"], "avg_launch_ms": fe["avg_ms"],
        "sms": rest_sms or 148,
        "sms_note": ("in the timed region this kernel runs on %d of the 148 SMs (the PLL's serial loops own the other %d, DESIGN.md 4.3.3) and shares them with "
                     "the next sub-chunks' prediction / table kernels; `achieved` is against the whole GPU's peak all the same - whole_job_launches has "
                     "the kernel alone on all SMs" % (rest_sms, loop_sms)) if rest_sms else None,
        "algorithmic_bytes": int(alg["frontend"]["bytes"] * pairs_per_step),
        "traffic": int(NCU_FRONTEND_DRAM_BYTES_PER_PAIR * pairs_per_step),
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per IQ pair from profiles/r2_ncu_k_frontend_stream.md (ncu --set full, whole-job launch), scaled to this launch",
        "whole_job_launches": isolated,
        "whole_job_note": "same kernels, one launch per step over the whole batch, not overlapped with the PLL (second, untimed-for-value pass)",
    }
    pll = kernels.get("pll")
    pll_info = None
    if pll:
        pll_ms_per_step = prof["pll"]["ms"] / args.steps           # the serial loops of all sub-chunks of a step
        pll_info = {"share_of_step": pll["share"], "avg_launch_ms": pll["avg_ms"], "ms_per_step": round(pll_ms_per_step, 3),
                    "ns_per_sample_per_stream": round(pll_ms_per_step * 1e6 / n_if, 2),
                    "stream_samples_per_s": round(S * n_if / (pll_ms_per_step * 1e-3), 0),
                    "loop": "table-driven (k_pll_predict -> k_pll_table_ops: exact two-candidate rows -> k_pll_sel: serial picks, certified per lane; DESIGN.md 4.3)" if PLL_TABLE else "direct (dy4_pllmath.h in the serial loop)",
                    "note": ("serial recurrence per stream, one warp per stream: the transcendental work runs beforehand in time-parallel kernels (counted in "
                             "pll_aux), the serial loop is (packed) float adds, one compare and selects - bound by the issue rate of a lone warp on that dependent "
                             "chain: 14.5 ns per sample with a sub-partition to itself, 17.8 two warps to a sub-partition (256 streams on the 32 SMs "
                             "set aside for the loops), 21-25 when FIR kernels share its SMs - not by FLOPs or bytes") if PLL_TABLE else
                            "serial recurrence per stream, one thread per stream: bound by the latency of its dependent FP64 chain, not by FLOPs or bytes"}

    # ---- CPU baseline: the reference's own code on the host cores, bounded sample (N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, kind, pairs, dt = cpu_throughput(8 * cores, BLOCKS_PER_STREAM, cores)
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d of the workload's streams x %d blocks (%.2f s of signal each), one reference pipeline per stream on %d worker "
                         "processes, %.1f s wall (%.0f core-seconds)" % (8 * cores, BLOCKS_PER_STREAM, BLOCKS_PER_STREAM * 0.021333, cores, dt, dt * cores)}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_max / args.steps, 3), "step_latency_ms": round(min(lat), 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(overlap),
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": S * nb * m.block_size,
                "d2h_bytes_per_step": S * n_audio * nch * 2, "steps": e2e_steps, "ms_per_step": round(e2e_ms / e2e_steps, 3),
                "timing": ("host clock around the process_host() calls (queued back to back: the upload of step k+1 runs beside the kernels of step k), the closing "
                           "dy4_pipeline_sync and the barrier, max over ranks") if overlap else
                          "host clock around the synchronous process_host() calls and the closing barrier, max over ranks",
                "pcm_equals_device_path": pcm_matches, "gather": gather_info},
        "oracle_check": oracle_check,
        "signal": "periodic over one step: the streams continue from step to step without a jump" if periodic else "the same buffer again every step: a phase jump per step",
        "gpu_launches": total_launches,
        "sm_partition": {"loop_sms": loop_sms, "other_sms": rest_sms} if loop_sms else None,
        "roofline": roofline, "kernels": kernels, "pll": pll_info, "cpu_baseline": cpu,
        "clocks": sampler.result(),
    }
    print(json.dumps(line))
    return 0


def main():
    global STREAMS_PER_GPU, BLOCKS_PER_STREAM, MODE, RDS, STEREO
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dy4", choices=["dy4", "reference"])
    ap.add_argument("--chunk-blocks", type=int, default=0, help="blocks per H2D chunk in the e2e leg (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="join every step to the stream before the next is queued (default: steps overlap on the device)")
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU (default: BASELINE configs[1], 256)")
    ap.add_argument("--blocks", type=int, default=BLOCKS_PER_STREAM, help="blocks per stream per step (default 47 = 1.003 s)")
    ap.add_argument("--mode", type=int, default=MODE, help="receiver mode 0..3 (default 0)")
    ap.add_argument("--mono", action="store_true", help="mono receiver (BASELINE configs[0]'s path: no pilot / PLL / L-R branch)")
    ap.add_argument("--rds", action="store_true", help="BASELINE configs[3]: also run the RDS path (mode 0 stereo; e.g. --streams 4096 --blocks 6)")
    args = ap.parse_args()
    STREAMS_PER_GPU, BLOCKS_PER_STREAM, MODE, RDS = args.streams, args.blocks, args.mode, args.rds
    STEREO = 0 if args.mono else 1
    args.warmup = max(args.warmup, 3) if args.impl == "dy4" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
