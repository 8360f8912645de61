#!/usr/bin/env python3
"""bench.py — ORB extract + Hamming match throughput at 1280x720 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

A "step" is one pass of the hot path over one batch of B synthetic 1280x720 RGB-D frames:
8-level pyramid -> per-cell FAST -> quadtree -> orientation -> blur -> rBRIEF-256 -> depth filter ->
brute-force Hamming match of every frame against its predecessor with `distance < 50`
(BASELINE.json configs[1]; reference frontend.cpp:1094-1132).
  value : frames/s with the batch resident in HBM (orbx_track_batch_device), whole job over all ranks
  e2e   : frames/s through the host-buffer C-ABI call (orbx_track_batch_submit / orbx_batch_wait): pinned host frames in,
          keypoints/descriptors/matches out, H2D and D2H inside the timed region; `link_ceiling_gbs` = an H2D-only probe of the
          same bytes run concurrently on all ranks
  parity_in_bench : the first S frames of the timed GPU batch compared with the reference CPU path's outputs for the same frames
Other BASELINE configs ride along as their own objects: config0 (640x480 pair, k=2 + ratio), config2 (one 4096-frame job,
block-partitioned over the ranks with preamble frames), config4 (per-frame YOLO boxes), association (configs[3], sharded DB + NCCL).
The CPU arm is the reference's ORBextractor.cpp itself (oracle/_ref, compiled unmodified; kind "reference") where that library
exists, else the C restatement (kind "port").  PyTorch is plumbing only (device buffers, CUDA events, torch.distributed).
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries the ONE JSON line and nothing else: libraries that chat on file descriptor 1 (NCCL prints its version banner there) are
# pointed at stderr for the whole run, the line itself goes to a duplicate of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

W, H = 1280, 720
SEED = 20261018
WORKLOAD = ("configs[1]: 1280x720 RGB-D stream, 8-level pyramid scale 1.2, depth-filtered extraction + "
            "frame-to-frame matching (k=1, distance<50)")
METRIC = "frames/s ORB extract+Hamming match, 1280x720, 1/2/4/8 B200; p50 latency"       # BASELINE.json "metric", verbatim
# algorithmic bytes per 1280x720 frame, staged model; pixels summed over the 8 levels
LEVEL_PX = [1280 * 720, 1067 * 600, 889 * 500, 741 * 417, 617 * 347, 514 * 289, 429 * 241, 357 * 201]
ALG_BYTES = {
    "k_resize_linear": sum(LEVEL_PX[:7]) + sum(LEVEL_PX[1:]),      # read L0..L6 once, write L1..L7 once (7 launches)
    "k_fast_cells": sum(LEVEL_PX),                                  # read every level once
    "k_fast_dense": sum(LEVEL_PX),                                  # the same levels, whole-level tiles (its score map is scratch, not algorithmic traffic)
    "k_blur7": 2 * sum(LEVEL_PX),                                   # read + write every level once
}
POPC_PER_PAIR = 5.0              # csrc/orbx_hamming.h (ORBX_MATCH_CSA = 2)
CAP = 1280


def _newest_traffic():
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    try:
        return json.load(open(files[-1])), os.path.basename(files[-1])
    except Exception:
        return None, None


def ncu_traffic(kernel, frames, launches_per_step):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full summary (profiles/*_traffic.json), scaled
    from the capture's batch to this run's; None when no capture is committed."""
    t, src = _newest_traffic()
    t = (t or {}).get(kernel)
    return (t["dram_bytes_per_frame_per_step"] * frames / max(1.0, launches_per_step), src) if t else (None, None)


def ncu_warp_inst(kernel, frames):
    t, _ = _newest_traffic()
    t = (t or {}).get(kernel)
    return t["warp_inst_per_frame_per_step"] * frames if t and t.get("warp_inst_per_frame_per_step") else None


def bind_to_gpu_numa(local_rank):
    """Multi-rank runs: pin this process to the CPUs nvidia-smi reports as local to its GPU, so that the pinned staging buffers are
    first-touched on that NUMA node and the H2D/D2H DMA does not cross sockets.  Best effort; returns the CPU list or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            cols = line.split()
            if cols and cols[0] == "GPU%d" % local_rank:
                for c in cols[1:]:
                    if c[0].isdigit() and ("-" in c or "," in c) and not c.startswith("NV"):
                        cpus = set()
                        for part in c.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                        os.sched_setaffinity(0, cpus)
                        return c
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------------
# the reference's CPU path
# ------------------------------------------------------------------------------------------------------------------------------
class CpuPath:
    """extract -> filterDepth -> match vs previous + `distance < 50` (reference frontend.cpp:1094-1132) on the host cores.
    Extraction = the reference's own ORBextractor.cpp (oracle/_ref, one extractor per OpenMP thread) when that library is present,
    else the C restatement; filterDepth and BFMatcher are the C restatements (the reference calls OpenCV for the latter)."""

    def __init__(self):
        import c_oracle as co
        co.build()
        self.co = co
        self.ro = None
        try:
            import ref_oracle as ro
            if ro.available():
                ro.lib()
                self.ro = ro
        except Exception:
            self.ro = None
        self.kind = "reference" if self.ro is not None else "port"
        self.orc = co.COracle()

    def extract(self, frames, nthreads):
        if self.ro is not None:
            return self.ro.extract_batch(frames, cap=2048, nthreads=nthreads)
        return self.orc.extract_batch(frames, cap=2048, nthreads=nthreads)

    def run(self, frames, depths, nthreads, boxes=None, keep=False):
        co = self.co
        t0 = time.perf_counter()
        kps, desc, counts = self.extract(frames, nthreads)
        prev, nmatch, out = None, 0, []
        for f in range(len(frames)):
            k, d = kps[f, :counts[f]], desc[f, :counts[f]]
            if depths is not None:
                k, d, _ = co.filter_depth(k, d, depths[f])
            if boxes is not None:
                k, d = co.filter_boxes(k, d, boxes[f], 1)
            m = co.match(d, prev, nthreads=nthreads) if prev is not None and len(d) and len(prev) else np.zeros(0, co.DM_DTYPE)
            good = m[m["distance"] < 50.0]
            nmatch += len(good)
            if keep:
                out.append((k.copy(), d.copy(), good.copy()))
            prev = d
        return time.perf_counter() - t0, nmatch, out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path with all host threads (bounded sample per step)."""
    if rank != 0:
        return
    cp = CpuPath()
    co = cp.co
    cores = os.cpu_count() or 1
    S = max(8, min(32, cores))                      # bounded sample of the 1280x720 stream per step
    frames = np.stack([co.synth_gray(SEED, f, W, H) for f in range(S)])
    depths = np.stack([co.synth_depth(SEED, f, W, H) for f in range(S)])
    for _ in range(args.warmup):
        cp.run(frames, depths, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cp.run(frames, depths, cores)[0]
    fps = S * args.steps / t
    sample = "%d frames/step of the synthetic 1280x720 RGB-D stream, extract+filterDepth+match; extraction = %s" % (
        S, "the reference's ORBextractor.cpp compiled unmodified (oracle/_ref), one instance per OpenMP thread" if cp.kind == "reference" else "C restatement (oracle port)")
    emit({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": S, "width": W, "height": H, "nfeatures": 1000, "nlevels": 8},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cp.kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def compare_frames(gpu, cpu, first_has_prev):
    """gpu: (kps [n,CAP] KP_DTYPE, desc [n,CAP,32], counts, matches [n,CAP] DM_DTYPE, mcounts); cpu: list of (kps, desc, good).
    Returns the number of frames that differ in keypoints, descriptors or (where the predecessor is known) matches."""
    kk, dd, cc, mm, mc = gpu
    bad = 0
    for f, (k, d, good) in enumerate(cpu):
        ok = cc[f] == len(k) and np.array_equal(kk[f, :cc[f]].view(np.uint8), k.view(np.uint8)) and np.array_equal(dd[f, :cc[f]], d)
        if ok and (f > 0 or first_has_prev):
            ok = mc[f] == len(good) and np.array_equal(mm[f, :mc[f]].view(np.uint8), good.view(np.uint8))
        bad += 0 if ok else 1
    return bad


# ------------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="frames resident per step and per GPU")
    ap.add_argument("--separate-blur", action="store_true", help="blur every level with its own kernel (ORBX_OPT_FUSED_BLUR = 0) instead of inside the descriptor kernel")
    ap.add_argument("--no-pdl", action="store_true", help="plain stream order instead of programmatic dependent launch (ORBX_OPT_PDL = 0)")
    ap.add_argument("--fast-ctas", type=int, default=0, help="resident FAST warps per SM in the overlapped schedule (0 = library default)")
    ap.add_argument("--overlap", action="store_true", help="two staggered half-batches on two streams (ORBX_OPT_OVERLAP = 1; measured slower than one chain)")
    ap.add_argument("--describe-all", action="store_true", help="the reference's order: describe every selected keypoint, then filter (ORBX_OPT_FILTER_FIRST = 0)")
    ap.add_argument("--fast-dense", type=int, default=0, help="ORBX_OPT_FAST_DENSE: 1 = the dense FAST formulation for batches, 3 = the same with the NMS inside the tile kernel")
    ap.add_argument("--match-engine", type=int, default=-1, help="ORBX_OPT_MATCH_MMA: 1 = default (tensor-memory kernel for large calls), 2 = the mma.sync kernel, 3 = tensor-memory kernel always, 0 = POPC")
    ap.add_argument("--popc-match", action="store_true", help="the LOP3/POPC matcher instead of the int8 tensor-core GEMM (ORBX_OPT_MATCH_MMA = 0)")
    ap.add_argument("--host-chunk", type=int, default=0, help="frames per pipeline chunk of the host-buffer call (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-assoc", action="store_true", help="skip the landmark-association leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the config0 / config2 / config4 legs")
    ap.add_argument("--kernels-only", action="store_true", help="device-resident leg only (for ncu captures)")
    ap.add_argument("--job-frames", type=int, default=4096, help="frames of the configs[2] job (whole job, split over the ranks)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import orbx
    from orbx import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ORB path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B, K = args.batch, args.steps
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, device=local_rank, max_keypoints=CAP, host_chunk=args.host_chunk)
    if args.fast_ctas:
        ex.set_fast_ctas(args.fast_ctas)
    if args.separate_blur:
        ex.set_fused_blur(False)
    if args.no_pdl:
        ex.set_pdl(False)
    if args.overlap:
        ex.set_overlap(True)
    if args.describe_all:
        ex.set_filter_first(False)
    if args.popc_match:
        ex.set_match_mma(False)
    elif args.match_engine >= 0:
        ex.set_match_mma(args.match_engine)
    if args.fast_dense:
        ex.set_fast_dense(args.fast_dense)
    L, hnd = ex.L, ex.handle
    stream = torch.cuda.ExternalStream(ex.stream, device=dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic inputs generated in HBM; rank r owns frames [r*B, (r+1)*B) of the stream ----
    gray = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    depth = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    ex._check(L.orbx_synth_gray_device(hnd, SEED, rank * B, B, W, H, gray.data_ptr(), W, W * H))
    ex._check(L.orbx_synth_depth_device(hnd, SEED, rank * B, B, W, H, depth.data_ptr(), 2 * W, 2 * W * H))
    kps = torch.empty((B, CAP, 28), dtype=torch.uint8, device=dev)
    desc = torch.empty((B, CAP, 32), dtype=torch.uint8, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    matches = torch.empty((B, CAP, 16), dtype=torch.uint8, device=dev)
    mcounts = torch.zeros(B, dtype=torch.int32, device=dev)

    def step():
        ex._check(L.orbx_track_batch_device(hnd, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                            kps.data_ptr(), desc.data_ptr(), CAP, counts.data_ptr(),
                                            matches.data_ptr(), mcounts.data_ptr(), ct.c_float(50.0)))

    def barrier():
        ex.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def gpu_outputs(n):
        return (kps[:n].cpu().numpy().view(orbx.KP_DTYPE).reshape(n, CAP), desc[:n].cpu().numpy(), counts[:n].cpu().numpy(),
                matches[:n].cpu().numpy().view(orbx.DM_DTYPE).reshape(n, CAP), mcounts[:n].cpu().numpy())

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ex.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    gpu_launches = ex.launch_count - launches0
    nkp = counts.cpu().numpy()
    nm = mcounts.cpu().numpy()
    S_par = min(B, 64 if world == 1 else 4)
    timed_out = gpu_outputs(S_par)                  # outputs of the LAST timed step (the same batch every step): checked against the CPU path below
    # per-kernel device time for the roofline: the same K steps again with every kernel on ONE stream (ORBX_OPT_SERIAL), so that
    # each CUDA-event bracket times its kernel alone
    ex.set_serial(True)
    step()
    barrier()
    ex.profile_enable(True)
    for _ in range(K):
        step()
    barrier()
    prof = ex.profile_read()
    ex.profile_enable(False)
    ex.set_serial(False)
    value = world * B * K / (ms * 1e-3)

    # ---- per-kernel device time and roofline of the dense stages ----
    hbm, peak_src = peaks()
    total_k = sum(v[0] for v in prof.values()) or 1.0
    kernels = {}
    for name, (kms, cnt) in prof.items():
        if cnt == 0:
            continue
        kernels[name] = {"ms_per_step": kms / K, "launches_per_step": cnt / K, "share": kms / total_k}
        if name in ALG_BYTES:
            gbs = ALG_BYTES[name] * B * K / (kms * 1e-3) / 1e9
            kernels[name].update({"achieved_gbs": gbs, "frac_of_hbm": gbs / hbm})
    dense = [n for n in ALG_BYTES if n in kernels]
    dom = max(dense, key=lambda n: kernels[n]["ms_per_step"]) if dense else None
    roofline = None
    if dom:
        traffic, traffic_src = ncu_traffic(dom, B, kernels[dom]["launches_per_step"])
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                    "frac": kernels[dom]["frac_of_hbm"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALG_BYTES[dom] * B / max(1.0, kernels[dom]["launches_per_step"]),
                    "dominant_overall": max(kernels, key=lambda n: kernels[n]["ms_per_step"]),
                    "step_frac": sum(ALG_BYTES.values()) * B * world * K / (ms * 1e-3) / 1e9 / hbm / world,
                    "note": "the kernel is an exact 8/16-bit integer pipeline bound by instruction issue, not by HBM: see `issue` (executed warp "
                            "instructions / kernel time against 4 schedulers x SMs x SM clock, per pipe in profiles/); its DRAM traffic equals its algorithmic bytes"}

    def issue_view(clk_mhz, sms):
        out = {}
        for name in dense:
            wi = ncu_warp_inst(name, B)
            if wi and clk_mhz:
                peak = 4.0 * sms * clk_mhz * 1e6
                ach = wi / (kernels[name]["ms_per_step"] * 1e-3)
                out[name] = {"warp_inst_per_step": wi, "thread_inst_per_pixel": wi * 32.0 / (B * sum(LEVEL_PX)) if name == "k_fast_cells" else None,
                             "achieved_ginst_s": ach / 1e9, "peak_ginst_s": peak / 1e9, "frac": ach / peak}
        return out or None

    match_roofline = None
    if "k_match_partial" in kernels:
        pairs = float((nkp[1:].astype(np.float64) * nkp[:-1]).sum() + float(nkp[0]) * float(nkp[-1]))      # frame f vs f-1; frame 0 vs the carried last frame
        popc = ex.bench_popc()
        t_s = kernels["k_match_partial"]["ms_per_step"] * 1e-3
        eng = 0 if args.popc_match else (args.match_engine if args.match_engine >= 0 else 1)
        match_roofline = {"kernel": {0: "k_match_partial", 2: "k_match_mma"}.get(eng, "k_match_umma"),
                          "engine": {0: "LOP3/POPC (5 POPC per pair after carry-save adders)",
                                     2: "int8 tensor-core GEMM (mma.sync m16n8k32.u8 on descriptors unpacked to 0/1 bytes, d = |q| + |t| - 2 q.t)"}.get(
                                     eng, "int8 GEMM on tcgen05 (tcgen05.mma kind::i8, 128 x 128 x 256 tiles, accumulators in TMEM; d = |q| + |t| - 2 q.t on descriptors unpacked to 0/1 bytes)"),
                          "bound": {0: "integer issue (ALU pipe)", 2: "tensor pipe (legacy mma.sync int8: ncu 58 % busy at 1.28 T pairs/s)"}.get(
                                   eng, "its own per-tile lock-step (ncu: tensor pipe 29 % busy, 30 % of the samples at the CTA barrier); the tensor pipe would allow ~9 T pairs/s"),
                          "pairs_per_step": pairs, "gpairs_per_s": pairs / t_s / 1e9,
                          "popc_equivalent": {"popc_per_pair": POPC_PER_PAIR, "achieved": POPC_PER_PAIR * pairs / t_s, "peak": popc, "unit": "POPC/s",
                                              "frac": POPC_PER_PAIR * pairs / t_s / popc if popc > 0 else None, "peak_source": "measured (orbx_bench_popc)",
                                              "note": "what the POPC kernel would need to sustain for the same pairs/s"}}
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    if args.kernels_only:
        clocks = sampler.stop()
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
                  "ms_per_step": ms / K, "kernels": kernels, "roofline": roofline, "match_roofline": match_roofline,
                  "issue": issue_view((clocks or {}).get("sm_mhz"), sms), "gpu_launches": int(gpu_launches), "clocks": clocks})
        return

    # ---- e2e: the host-buffer C-ABI calls, pinned host memory, H2D + D2H inside the timed region ----
    Be = B
    nbytes_g, nbytes_d = Be * W * H, Be * W * H * 2
    hp = {}
    sizes = {"gray": nbytes_g, "depth": nbytes_d}
    for j in range(2):
        sizes.update({"kps%d" % j: Be * CAP * 28, "desc%d" % j: Be * CAP * 32, "cnt%d" % j: Be * 4, "m%d" % j: Be * CAP * 16, "mc%d" % j: Be * 4})
    for k, n in sizes.items():
        hp[k] = L.orbx_alloc_pinned(n)
        if not hp[k]:
            raise SystemExit("pinned allocation failed")
    ex._check(L.orbx_copy_to_host(hnd, hp["gray"], gray.data_ptr(), nbytes_g))
    ex._check(L.orbx_copy_to_host(hnd, hp["depth"], depth.data_ptr(), nbytes_d))

    def step_e2e(nf=Be):                                            # synchronous call (latency leg)
        ex._check(L.orbx_track_batch(hnd, hp["gray"], nf, W, H, W, hp["depth"], 2 * W, hp["kps0"], hp["desc0"], CAP, hp["cnt0"],
                                     hp["m0"], hp["mc0"], ct.c_float(50.0)))

    def submit(j):
        t = ct.c_int32()
        ex._check(L.orbx_track_batch_submit(hnd, hp["gray"], Be, W, H, W, hp["depth"], 2 * W, hp["kps%d" % j], hp["desc%d" % j], CAP,
                                            hp["cnt%d" % j], hp["m%d" % j], hp["mc%d" % j], ct.c_float(50.0), ct.byref(t)))
        return t.value

    def run_async(n):
        prev = None
        for k in range(n):
            t = submit(k & 1)
            if prev is not None:
                ex._check(L.orbx_batch_wait(hnd, prev))
            prev = t
        ex._check(L.orbx_batch_wait(hnd, prev))

    run_async(args.warmup)
    barrier()
    t0 = time.perf_counter()
    run_async(K)                                                     # returns when the last step's results are in host memory
    wall = time.perf_counter() - t0
    barrier()
    ms_e = max_over_ranks(wall * 1e3)                                # host wall clock around submit..wait of K steps (copies are on other streams)
    for _ in range(2):
        step_e2e()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    wall_sync = time.perf_counter() - t0
    # link ceiling: the SAME gray bytes, H2D only, from the same pinned buffer, all ranks at once — what the host fabric gives this rank
    # while its neighbours pull too.  e2e frames/s x bytes per frame / this = how much of the link the pipeline uses.
    barrier()
    for _ in range(2):
        ex._check(L.orbx_copy_to_device(hnd, gray.data_ptr(), hp["gray"], nbytes_g))
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(4, K // 2)):
        ex._check(L.orbx_copy_to_device(hnd, gray.data_ptr(), hp["gray"], nbytes_g))
    link_s = (time.perf_counter() - t0) / max(4, K // 2)
    barrier()
    link_gbs_rank = nbytes_g / link_s / 1e9
    link_gbs_min = -max_over_ranks(-link_gbs_rank)
    link_gbs_sum = sum_over_ranks(link_gbs_rank)
    e2e_value = world * Be * K / (ms_e * 1e-3)
    e2e = {"value": e2e_value, "unit": "frames/s",
           # gray frames are DMA'd; the pinned depth maps are NOT copied: the depth filter gathers one 32-byte PCIe sector per
           # selected keypoint in place (<= max_keypoints per frame) — counted here at that upper bound
           "h2d_bytes_per_step": nbytes_g + Be * CAP * 32, "depth_bytes_resident_on_host": nbytes_d,
           "d2h_bytes_per_step": Be * (CAP * (28 + 32 + 16) + 8),
           "api": "orbx_track_batch_submit + orbx_batch_wait, two batches in flight, host pinned buffers, zero-copy depth gather",
           "ms_per_step": ms_e / K, "timing": "host wall clock around K submit/wait steps (work spans three streams)",
           "sync_call_frames_per_s": world * Be * K / wall_sync, "sync_call_api": "orbx_track_batch (one blocking call per step, chunk pipeline inside)",
           "link_ceiling_gbs": link_gbs_min, "link_ceiling_gbs_all_ranks": link_gbs_sum,
           "link_probe": "H2D of the same %d MB pinned gray buffer, all %d ranks concurrently, no kernels" % (nbytes_g // 1000000, world),
           "frac_of_link": (e2e_value / world) * (W * H) / 1e9 / link_gbs_min if link_gbs_min > 0 else None}

    clocks = sampler.stop()          # sampled from the first timed step to the end of the e2e leg (the GPU is under load throughout)
    # ---- per-frame latency, batch = 1 through the same host call (configs[1] p50) ----
    lat = []
    ex.track_reset()
    for i in range(220):
        t0 = time.perf_counter()
        step_e2e(1)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat[20:])
    latency = {"p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)), "frames": len(lat),
               "api": "orbx_track_batch, nframes=1, host buffers"}

    cp = None
    if rank == 0:
        cp = CpuPath()
    co = cp.co if cp else None

    # ---- landmark association (configs[3]): 2048 queries vs a 1M-row database sharded over the ranks, NCCL all-gather inside the C ABI ----
    assoc = None
    if not args.no_assoc:
        import c_oracle as co_all
        co_all.build()
        NQ, ROWS = 2048, 1 << 20
        first_r, rows_r = sharding.block_range(ROWS, world, rank)
        db = orbx.LandmarkDB(ex, rows_r, first_index=first_r)
        drows = torch.empty((rows_r, 32), dtype=torch.uint8, device=dev)
        ex._check(L.orbx_synth_descriptors_device(hnd, 1234, first_r, rows_r, drows.data_ptr()))
        db.append_device(drows.data_ptr(), rows_r)
        # queries: DB rows i*512 with 0..40 random bit flips (true neighbour below 50, spread over every shard) + 10 % random rows
        rng = np.random.default_rng(4321)
        qh = np.zeros((NQ, 32), np.uint8)
        n_true = NQ - NQ // 10
        for i in range(n_true):
            row = co_all.synth_descriptors(1234, (i * 512) % ROWS, 1)[0].copy()
            for b in rng.choice(256, size=int(rng.integers(0, 41)), replace=False):
                row[b >> 3] ^= np.uint8(1 << (b & 7))
            qh[i] = row
        qh[n_true:] = rng.integers(0, 256, (NQ - n_true, 32), dtype=np.uint8)
        q = torch.from_numpy(qh).to(dev)
        merged = torch.empty((NQ, 4), dtype=torch.int32, device=dev)
        comm = orbx.Comm(ex, dist=dist) if dist is not None else orbx.Comm(ex, nranks=1, rank=0, unique_id=_single_rank_id(L))

        def assoc_step():
            db.query_top2_sharded_device(comm, q.data_ptr(), NQ, merged.data_ptr())      # per-shard kernel -> ncclAllGather -> merge, one stream, no host sync

        def time_assoc():
            for _ in range(3):
                assoc_step()
            barrier()
            ex.profile_enable(True)
            e0.record(stream)
            for _ in range(K):
                assoc_step()
            e1.record(stream)
            barrier()
            pr = ex.profile_read()
            ex.profile_enable(False)
            return max_over_ranks(e0.elapsed_time(e1) / K), pr

        peer = comm.peer_memory
        comm.set_transport(nccl=True)
        ms_nccl, _ = time_assoc()
        merged_nccl = merged.clone()
        comm.set_transport(nccl=False)
        ms_a, profa = time_assoc()                  # default transport: peer-memory mailboxes where the ranks could map each other, else NCCL again
        same_transports = bool((merged_nccl == merged).all().item())
        res = merged.cpu().numpy().view(np.uint32)
        # the merged answer against ONE unsharded database on this GPU (every rank checks its own copy of the merged result)
        full = orbx.LandmarkDB(ex, ROWS, first_index=0)
        frows = torch.empty((ROWS, 32), dtype=torch.uint8, device=dev)
        ex._check(L.orbx_synth_descriptors_device(hnd, 1234, 0, ROWS, frows.data_ptr()))
        full.append_device(frows.data_ptr(), ROWS)
        single = torch.empty((NQ, 4), dtype=torch.int32, device=dev)
        full.query_top2_device(q.data_ptr(), NQ, single.data_ptr())
        ex.sync()
        mism = int(sum_over_ranks(float((single != merged).any(dim=1).sum().item())))
        cpu_check = None
        if rank == 0:
            QC = 48
            want = co.knn2(qh[:QC], frows.cpu().numpy())
            got = res[:QC]
            okc = (got[:, 0] == want[:, 0]["distance"].astype(np.uint32)) & (got[:, 1] == want[:, 0]["trainIdx"].astype(np.uint32)) & \
                  (got[:, 2] == want[:, 1]["distance"].astype(np.uint32)) & (got[:, 3] == want[:, 1]["trainIdx"].astype(np.uint32))
            cpu_check = {"queries": QC, "mismatches": int((~okc).sum()), "against": "BFMatcher knnMatch(k=2) restatement (oracle) over the full 1M-row database"}
        popc = ex.bench_popc()
        kms = profa["k_match_partial"][0] / max(1, profa["k_match_partial"][1])
        assoc = {"queries": NQ, "db_rows": ROWS, "rows_per_gpu": rows_r, "ms_per_query_batch": ms_a,
                 "gpairs_per_s": NQ * ROWS / (ms_a * 1e-3) / 1e9,
                 "query_set": "%d database rows i*512 with 0..40 bit flips + %d random rows" % (n_true, NQ - n_true),
                 "nn_below_50": int((res[:, 0] < 50).sum()), "nn_from_other_shards": int((res[:, 1] >= rows_r).sum()) if world > 1 else 0,
                 "mismatches_vs_unsharded": mism, "cpu_check": cpu_check,
                 "kernel_ms": kms, "kernel_gpairs_per_s": NQ * rows_r / (kms * 1e-3) / 1e9 if kms > 0 else None,
                 "engine": "k_match_partial (POPC)" if args.popc_match else ("k_match_mma (mma.sync int8 GEMM)" if args.match_engine == 2 else "k_match_umma (tcgen05.mma kind::i8, accumulators in TMEM)"), "popc_per_s_measured_peak": popc,
                 "transport": "peer memory (NVLink stores into every peer's mailbox + sequence flags, one kernel)" if peer else "nccl",
                 "ms_per_query_batch_nccl": ms_nccl, "transports_agree": same_transports,
                 "collective": "orbx_db_query_top2_sharded_device: per-shard kernel -> exchange of 32 KB/rank -> merge kernel, all on the handle's stream, "
                               "no host synchronisation; timed with CUDA events on that stream; exchange = peer-memory mailboxes (default) or ncclAllGather"}
        # ---- the whole associateObservation call (descriptor gate + reprojection gate, backend.cpp:1064-1120) against the sharded database ----
        prng = np.random.default_rng(99)
        pos_all = np.stack([prng.uniform(-2, 2, ROWS), prng.uniform(-1.2, 1.2, ROWS), prng.uniform(0.4, 6.0, ROWS)], 1).astype(np.float32)
        Rm, tv, Kc = np.eye(3), np.zeros(3), (615.3, 615.9, 640.2, 360.4)
        pose = orbx.LandmarkDB.pose(Rm, tv, *Kc)
        db.set_positions(pos_all[first_r:first_r + rows_r])
        full.set_positions(pos_all)
        src = (np.arange(n_true) * 512) % ROWS
        uv = np.stack([Kc[0] * pos_all[src, 0].astype(np.float64) / pos_all[src, 2] + Kc[2], Kc[1] * pos_all[src, 1].astype(np.float64) / pos_all[src, 2] + Kc[3]], 1)
        qpx_h = np.zeros((NQ, 2), np.float32)
        qpx_h[:n_true] = (uv + prng.normal(0, 2.5, (n_true, 2))).astype(np.float32)          # some beyond the 5-px gate
        qpx_h[n_true:] = prng.uniform(0, 700, (NQ - n_true, 2)).astype(np.float32)
        qpx = torch.from_numpy(qpx_h).to(dev)
        g_out = torch.zeros((NQ, 16), dtype=torch.uint8, device=dev)

        def gated_step():
            db.associate_sharded_device(comm, q.data_ptr(), qpx.data_ptr(), NQ, pose, g_out.data_ptr())

        for _ in range(3):
            gated_step()
        barrier()
        e0.record(stream)
        for _ in range(K):
            gated_step()
        e1.record(stream)
        barrier()
        ms_g = max_over_ranks(e0.elapsed_time(e1) / K)
        got_g = g_out.cpu().numpy().view(orbx.ASSOC_DTYPE).reshape(-1)
        one = full.associate(qh, qpx_h, pose)                                                  # ONE unsharded database on this GPU
        mism_g = int(sum_over_ranks(float((got_g.view(np.uint8).reshape(NQ, 16) != one.view(np.uint8).reshape(NQ, 16)).any(axis=1).sum())))
        gated_cpu = None
        if rank == 0:
            QC = 48
            widx, werr, wdist = co_all.associate(qh[:QC], qpx_h[:QC], frows.cpu().numpy(), pos_all, Rm, tv, *Kc)
            okg = (got_g["landmark"][:QC] == widx) & ((widx < 0) | ((got_g["reproj_error"][:QC].view(np.uint64) == werr.view(np.uint64)) & (got_g["distance"][:QC] == wdist)))
            gated_cpu = {"queries": QC, "mismatches": int((~okg).sum()), "against": "associateObservation restatement (oracle) over the full 1M-row database"}
        assoc["gated"] = {"call": "orbx_db_associate_sharded_device (Hamming < 50, then smallest reprojection error < 5 px): per-shard kernel -> exchange of 32 KB/rank -> merge",
                          "ms_per_query_batch": ms_g, "gpairs_per_s": NQ * ROWS / (ms_g * 1e-3) / 1e9, "associated": int((got_g["landmark"] >= 0).sum()),
                          "mismatches_vs_unsharded": mism_g, "cpu_check": gated_cpu,
                          "engine": "k_assoc_partial (POPC)" if args.popc_match else ("k_assoc_mma (mma.sync int8 GEMM, reprojection in the epilogue)" if args.match_engine == 2 else "k_assoc_umma (tcgen05 int8 GEMM, accumulators in TMEM, reprojection in the epilogue)")}
        comm.close(); db.close(); full.close()
        del drows, frows

    extra = {}
    if not args.no_configs:
        extra["config0"] = leg_config0(ex, orbx, cp, rank)
        extra["config2"] = leg_config2(ex, orbx, torch, dist, dev, stream, rank, world, args.job_frames, B, cp, max_over_ranks, sum_over_ranks, barrier)
        extra["config4"] = leg_config4(ex, orbx, torch, dev, stream, rank, world, B, K, gray, depth, kps, desc, counts, matches, mcounts, cp, max_over_ranks, barrier)

    # ---- reference CPU path timed beside it (rank 0, N = 1 only) + parity of the timed GPU batch against it ----
    cpu_baseline, parity = None, None
    if rank == 0:
        cores = os.cpu_count() or 1
        if world == 1 and not args.no_cpu:
            S = max(16, min(64, 2 * cores))
            fr = np.stack([co.synth_gray(SEED, f, W, H) for f in range(S)])
            dp = np.stack([co.synth_depth(SEED, f, W, H) for f in range(S)])
            cp.run(fr[:4], dp[:4], cores)
            dt1, _, kept = cp.run(fr, dp, cores, keep=True)
            REP = int(max(2, min(200, np.ceil(10.0 / max(dt1, 1e-3)))))          # a bounded sample of about 10 s of wall time on all cores
            dt = sum(cp.run(fr, dp, cores)[0] for _ in range(REP))
            cpu_baseline = {"value": S * REP / dt, "unit": "frames/s", "cores": cores, "kind": cp.kind,
                            "sample": "%d passes over %d frames of the same synthetic stream (%.1f s), extract+filterDepth+match, OpenMP over %d threads; extraction = %s"
                                      % (REP, S, dt, cores, "the reference's ORBextractor.cpp compiled unmodified (oracle/_ref)" if cp.kind == "reference" else "C restatement")}
        else:
            S = S_par
            fr = np.stack([co.synth_gray(SEED, f, W, H) for f in range(S)])
            dp = np.stack([co.synth_depth(SEED, f, W, H) for f in range(S)])
            _, _, kept = cp.run(fr, dp, cores, keep=True)
        n = min(S_par, len(kept))
        parity = {"frames": n, "mismatches": compare_frames(tuple(a[:n] for a in timed_out), kept[:n], first_has_prev=False),
                  "checked": "keypoints (28-byte records, order), descriptors, matches<50 of frames 1.. of the LAST timed step's batch (orbx_track_batch_device, batch %d) "
                             "against the CPU path (%s)" % (B, cp.kind)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "width": W, "height": H,
                       "nfeatures": 1000, "nlevels": 8, "l2_policy": "inputs larger than L2 (%.0f MB per step per GPU)" % (B * W * H * 3 / 1e6),
                       "sharding": "frame-parallel, no data-path collective", "cpu_affinity": numa,
                       "schedule": "one dependent chain pyramid -> FAST -> quadtree -> describe (7x7 Gaussian evaluated inside, at the sample points) -> filter -> match; `kernels`/`roofline` timed in a second pass of the same K steps with every kernel on one stream (ORBX_OPT_SERIAL)"},
            "roofline": roofline, "match_roofline": match_roofline,
            "issue": issue_view((clocks or {}).get("sm_mhz"), sms), "kernels": kernels, "cpu_baseline": cpu_baseline, "e2e": e2e, "latency": latency,
            "parity_in_bench": parity, "association": assoc, "gpu_launches": int(gpu_launches), "clocks": clocks,
            "keypoints_per_frame": float(nkp.mean()), "matches_per_frame": float(nm.mean()),
        }
        out.update(extra)
        emit(out)
    for p in hp.values():
        L.orbx_free_pinned(p)
    if dist is not None:
        dist.destroy_process_group()


def _single_rank_id(L):
    raw = (ct.c_uint8 * 128)()
    if L.orbx_comm_get_unique_id(raw) != 0:
        raise SystemExit("NCCL unavailable: " + (L.orbx_comm_last_error() or b"").decode())
    return bytes(raw)


# ------------------------------------------------------------------------------------------------------------------------------
def leg_config0(ex, orbx, cp, rank):
    """configs[0]: one synthetic 640x480 frame pair, 1000 features, kNN k=2 + ratio 0.75 (exact in integers: 4*d0 < 3*d1).
    GPU: two extractions + orbx_match(k=2, ratio) through the host API.  CPU beside it: cv::ORB + BFMatcher.knnMatch (the reference's
    gtest profile, cv2) at 1 thread and at all cores, and the frontend's own extractor (oracle/_ref) + the knn restatement."""
    if rank != 0:
        return None
    co = cp.co
    w, h = 640, 480
    g0, g1 = co.synth_gray(SEED, 0, w, h), co.synth_gray(SEED, 1, w, h)
    e = orbx.ORBextractor(max_width=w, max_height=h, device=ex.params.device)
    try:
        def pair():
            k0, d0 = e(g0)
            k1, d1 = e(g1)
            return k0, d0, k1, d1, e.match(d1, d0, k=2, ratio=0.75)
        for _ in range(5):
            out = pair()
        ts = []
        for _ in range(40):
            t0 = time.perf_counter()
            out = pair()
            ts.append((time.perf_counter() - t0) * 1e3)
        k0, d0, k1, d1, good = out
    finally:
        e.close()
    # the same pair through profile C (cv::ORB on the GPU: INTER_LINEAR_EXACT pyramid, whole-level FAST, Harris + retainBest, float blur)
    ec = orbx.ORBextractor(max_width=w, max_height=h, device=ex.params.device, profile="cvorb")
    try:
        def pair_c():
            a0, b0 = ec(g0)
            a1, b1 = ec(g1)
            return a0, b0, a1, b1, ec.match(b1, b0, k=2, ratio=0.75)
        for _ in range(5):
            outc = pair_c()
        tc = []
        for _ in range(40):
            t0 = time.perf_counter()
            outc = pair_c()
            tc.append((time.perf_counter() - t0) * 1e3)
    finally:
        ec.close()
    # parity of this leg: extraction against the CPU path, kNN + ratio against the restatement
    r0 = cp.extract(g0[None], 1)
    r1 = cp.extract(g1[None], 1)
    same = (np.array_equal(k0.view(np.uint8), r0[0][0, :r0[2][0]].view(np.uint8)) and np.array_equal(d0, r0[1][0, :r0[2][0]]) and
            np.array_equal(k1.view(np.uint8), r1[0][0, :r1[2][0]].view(np.uint8)) and np.array_equal(d1, r1[1][0, :r1[2][0]]))
    k2 = co.knn2(d1, d0)
    keep = (k2[:, 1]["trainIdx"] >= 0) & (4 * k2[:, 0]["distance"].astype(np.int64) < 3 * k2[:, 1]["distance"].astype(np.int64))
    same = bool(same and np.array_equal(good.view(np.uint8), k2[keep, 0].view(np.uint8)))
    res = {"workload": "configs[0]: 640x480 synthetic frame pair, 1000 features, kNN k=2 + ratio 0.75", "gpu_pair_ms_p50": float(np.percentile(ts, 50)),
           "gpu_api": "2 x orbx_extract + orbx_match(k=2, ratio=0.75), host buffers, blocking calls", "keypoints": [int(len(k0)), int(len(k1))],
           "ratio_matches": int(len(good)), "parity_mismatch": not same, "cores": os.cpu_count()}
    res["gpu_cvorb_pair_ms_p50"] = float(np.percentile(tc, 50))
    res["gpu_cvorb_keypoints"] = [int(len(outc[0])), int(len(outc[2]))]
    res["gpu_cvorb_ratio_matches"] = int(len(outc[4]))
    t_cpu = []
    for _ in range(5):
        t0 = time.perf_counter()
        a, b = cp.extract(g0[None], 1), cp.extract(g1[None], 1)
        co.knn2(b[1][0, :b[2][0]], a[1][0, :a[2][0]], nthreads=1)
        t_cpu.append((time.perf_counter() - t0) * 1e3)
    res["cpu_frontend_extractor_pair_ms_1thread"] = float(np.median(t_cpu))
    res["cpu_frontend_extractor_kind"] = cp.kind
    try:
        import cv2
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        for label, nt in (("1thread", 1), ("allcores", 0)):
            cv2.setNumThreads(nt if nt else (os.cpu_count() or 1))
            orb = cv2.ORB_create(1000)
            t_cv = []
            for _ in range(12):
                t0 = time.perf_counter()
                ka, da = orb.detectAndCompute(g0, None)
                kb, db_ = orb.detectAndCompute(g1, None)
                mm = bf.knnMatch(db_, da, k=2)
                _ = [m for m in mm if len(m) == 2 and m[0].distance < 0.75 * m[1].distance]
                t_cv.append((time.perf_counter() - t0) * 1e3)
            res["cv2_orb_knn_pair_ms_" + label] = float(np.median(t_cv[2:]))
            res["cv2_threads_" + label] = int(cv2.getNumThreads())
            if nt == 1:                                        # profile C on the GPU against cv::ORB itself: same keypoint sets, same ratio matches
                sa = {(p.octave, float(np.float32(p.pt[0])), float(np.float32(p.pt[1]))) for p in ka}
                ga = {(int(q["octave"]), float(q["x"]), float(q["y"])) for q in outc[0]}
                want = sorted((m[0].queryIdx, m[0].trainIdx) for m in bf.knnMatch(outc[3], outc[1], k=2) if len(m) == 2 and 4 * m[0].distance < 3 * m[1].distance)
                res["gpu_cvorb_vs_cv2"] = {"keypoint_set_equal": sa == ga, "ratio_matches_equal": want == sorted(zip(outc[4]["queryIdx"].tolist(), outc[4]["trainIdx"].tolist()))}
        res["cv2_version"] = cv2.__version__
    except Exception as exc:                                   # cv2 is optional on the GPU box
        res["cv2"] = "unavailable: %s" % type(exc).__name__
    return res


def leg_config2(ex, orbx, torch, dist, dev, stream, rank, world, job_frames, B, cp, max_over_ranks, sum_over_ranks, barrier):
    """configs[2]: ONE job of `job_frames` frames block-partitioned over the ranks (sharding.stream_block): every rank but the first
    re-extracts its predecessor frame as a one-frame preamble, so the pair that straddles two ranks is matched exactly as in the
    unsharded stream; no data-path collective.  Frames are pre-staged in HBM (gray only: extraction + frame-to-frame matching)."""
    from orbx import sharding
    L, hnd = ex.L, ex.handle
    first, count, pre = sharding.stream_block(job_frames, world, rank)
    gray = torch.empty((count, H, W), dtype=torch.uint8, device=dev)
    for c0 in range(0, count, 512):
        n = min(512, count - c0)
        ex._check(L.orbx_synth_gray_device(hnd, SEED, first + c0, n, W, H, gray[c0:].data_ptr(), W, W * H))
    pre_g = torch.empty((1, H, W), dtype=torch.uint8, device=dev)
    if pre is not None:
        ex._check(L.orbx_synth_gray_device(hnd, SEED, pre, 1, W, H, pre_g.data_ptr(), W, W * H))
    kps = torch.empty((B, CAP, 28), dtype=torch.uint8, device=dev)
    desc = torch.empty((B, CAP, 32), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(count + 1, dtype=torch.int32, device=dev)
    m = torch.empty((B, CAP, 16), dtype=torch.uint8, device=dev)
    mc = torch.zeros(count + 1, dtype=torch.int32, device=dev)

    def job():
        ex.track_reset()
        if pre is not None:                                       # preamble: carries frame first-1's descriptors into the handle's state
            ex._check(L.orbx_track_batch_device(hnd, pre_g.data_ptr(), 1, W, H, W, W * H, None, 0, 0, kps.data_ptr(), desc.data_ptr(), CAP,
                                                cnt[count:].data_ptr(), m.data_ptr(), mc[count:].data_ptr(), ct.c_float(50.0)))
        for c0 in range(0, count, B):
            n = min(B, count - c0)
            ex._check(L.orbx_track_batch_device(hnd, gray[c0:].data_ptr(), n, W, H, W, W * H, None, 0, 0, kps.data_ptr(), desc.data_ptr(), CAP,
                                                cnt[c0:].data_ptr(), m.data_ptr(), mc[c0:].data_ptr(), ct.c_float(50.0)))

    job()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    job()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    tot_kp = sum_over_ranks(float(cnt[:count].sum().item()))
    tot_m = sum_over_ranks(float(mc[:count].sum().item()))
    # boundary pairs on hardware: the match count of this rank's FIRST frame (vs the preamble) must equal what the unsharded stream gives;
    # rank 0 recomputes every boundary pair (f, f-1) on its own GPU and on the CPU path
    bm = torch.zeros(world, dtype=torch.float64, device=dev)
    bm[rank] = float(mc[0].item())
    if dist is not None:
        dist.all_reduce(bm)
    check = None
    if rank == 0:
        bad, pairs = 0, sharding.boundary_pairs(job_frames, world)
        co = cp.co
        for r, (f, fprev) in enumerate(pairs, start=1):
            fr = np.stack([co.synth_gray(SEED, fprev, W, H), co.synth_gray(SEED, f, W, H)])
            _, _, kept = cp.run(fr, None, os.cpu_count() or 1, keep=True)
            bad += 0 if int(bm[r].item()) == len(kept[1][2]) else 1
        check = {"boundary_pairs": len(pairs), "mismatches": bad, "against": "the CPU path on the two frames of each pair (match count of the straddling pair)"}
    del gray
    torch.cuda.empty_cache()
    return {"workload": "configs[2]: one job of %d synthetic 1280x720 frames, frame-parallel extraction + frame-to-frame matching, block-partitioned over %d GPU(s) "
                        "with one preamble frame per rank (sharding.stream_block), frames pre-staged in HBM, no collective" % (job_frames, world),
            "frames": job_frames, "frames_per_gpu": count, "ms": ms, "frames_per_s": job_frames / (ms * 1e-3), "scaling": "strong",
            "keypoints_total": tot_kp, "matches_total": tot_m, "boundary_check": check}


def leg_config4(ex, orbx, torch, dev, stream, rank, world, B, K, gray, depth, kps, desc, counts, matches, mcounts, cp, max_over_ranks, barrier):
    """configs[4]: the configs[1] stream with per-frame YOLO boxes (4 per frame, class 0 = "person" dropped after selection):
    orbx_track_batch_boxes_device.  Parity of the first frames against the CPU path with Backend::categorizeObservation's rule."""
    import c_oracle as co
    L, hnd = ex.L, ex.handle
    fb = [co.synth_boxes(SEED, rank * B + f, W, H) for f in range(B)]
    boxes, off = ex.pack_frame_boxes(fb)
    d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1).copy()).to(dev)
    d_off = torch.from_numpy(off).to(dev)

    def step():
        ex._check(L.orbx_track_batch_boxes_device(hnd, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                  d_boxes.data_ptr(), d_off.data_ptr(), len(boxes), ct.c_uint64(1),
                                                  kps.data_ptr(), desc.data_ptr(), CAP, counts.data_ptr(),
                                                  matches.data_ptr(), mcounts.data_ptr(), ct.c_float(50.0)))
    for _ in range(3):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    res = {"workload": "configs[4]: configs[1] stream + per-frame YOLO boxes (4 per frame, 'person' boxes excluded after selection)",
           "frames_per_s": world * B * K / (ms * 1e-3), "ms_per_step": ms / K, "frames_per_step_per_gpu": B,
           "keypoints_per_frame": float(counts.float().mean().item()), "scaling": "weak"}
    if rank == 0:
        S = min(B, 6)
        fr = np.stack([co.synth_gray(SEED, f, W, H) for f in range(S)])
        dp = np.stack([co.synth_depth(SEED, f, W, H) for f in range(S)])
        _, _, kept = cp.run(fr, dp, os.cpu_count() or 1, boxes=fb[:S], keep=True)
        gpu = (kps[:S].cpu().numpy().view(orbx.KP_DTYPE).reshape(S, CAP), desc[:S].cpu().numpy(), counts[:S].cpu().numpy(),
               matches[:S].cpu().numpy().view(orbx.DM_DTYPE).reshape(S, CAP), mcounts[:S].cpu().numpy())
        res["parity"] = {"frames": S, "mismatches": compare_frames(gpu, kept, first_has_prev=False)}
    return res


if __name__ == "__main__":
    main()
