#!/usr/bin/env python3
"""bench.py -- ORB front-end throughput on B200, one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Metric (BASELINE.json): ORB frames/s on 1241x376 frames, 2000 features, 8 levels.
Workload at every N = BASELINE config[1]: 64-frame synthetic stereo batches (32 L/R pairs) per GPU
per step; frames are independent, so ranks share nothing on the data path (weak scaling: each
rank extracts its own 64-frame batch; value = frames of all ranks / max-over-ranks time).
  value : device-resident -- the step's 64 frames are already in HBM (rotating over a pool of
          batches larger than the 126 MB L2), timed with CUDA events on the launching stream.
  e2e   : through the C ABI entry point a caller uses (orbx_extract_batch) with pinned HOST
          buffers: H2D of the frames and D2H of keypoints + descriptors inside the timed region.
The same line carries the Hamming kNN-2 figures (2000 x 100000, query-sharded over the ranks with
one all_gather of the 16-byte result records) under "matching".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"),):
    if _p not in sys.path:
        sys.path.insert(0, _p)

W, H, NFEAT, NLEVELS, BATCH = 1241, 376, 2000, 8, 64
NQ, NT = 2000, 100000
P_PYR = 1444097                      # sum of level pixels, SURVEY Appendix C
B_ALG = W * H + P_PYR + NFEAT * 60   # 2 030 713 algorithmic bytes / frame, SURVEY 8(d)
POOL = 6                             # distinct batches resident in HBM (6 x 30 MB inputs > L2)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML polled every 2 ms from a thread
    (nvidia-smi -lms as the fallback when NVML is not importable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nv, self.stop_flag, self.thread = None, False, None
        self.sm, self.reasons_seen, self.sm_max = [], set(), None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for n, bit in names.items():
                            if r & bit:
                                self.reasons_seen.add(n)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nv = N
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
            return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.sm_max,
                    "reasons": sorted(self.reasons_seen), "samples": len(self.sm), "source": "NVML polled every 2 ms over the timed regions"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU side: the reference's own orbextractor.cpp (oracle/_ref) when it was built, else the C port
# --------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orb_oracle_py as O
        self.O = O
        ref = os.path.join(ROOT, "oracle", "_ref", "liborbref.so")
        self.kind = "port"
        if os.path.exists(ref):
            try:
                R = C.CDLL(ref)

                class Cfg(C.Structure):
                    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
                R.orbref_create.restype = C.c_void_p
                R.orbref_create.argtypes = [C.POINTER(Cfg)]
                R.orbref_destroy.argtypes = [C.c_void_p]
                R.orbref_run.restype = C.c_int
                R.orbref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
                self.R, self.Cfg, self.kind = R, Cfg, "reference"
            except OSError:
                pass

    def extract_frames(self, frames, threads):
        """frames/s over `frames` with `threads` host threads, one extractor instance per thread."""
        O = self.O
        cap = NFEAT + 256
        n = len(frames)
        if self.kind == "reference":
            def work(t):
                h = self.R.orbref_create(C.byref(self.Cfg(NFEAT, 1.2, NLEVELS, 20, 7)))
                kps = np.zeros(cap, O.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
                tot = 0
                for i in range(t, n, threads):
                    f = frames[i]
                    tot += self.R.orbref_run(h, f.ctypes.data, W, H, f.strides[0], kps.ctypes.data, desc.ctypes.data, cap)
                self.R.orbref_destroy(h)
                return tot
            t0 = time.perf_counter()
            ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            [t.start() for t in ths]; [t.join() for t in ths]
            return n / (time.perf_counter() - t0)
        ex = O.Extractor(NFEAT, 1.2, NLEVELS)
        t0 = time.perf_counter()
        ex.extract_batch_mt(frames, threads)
        return n / (time.perf_counter() - t0)

    def knn2(self, q, t, threads):
        t0 = time.perf_counter()
        self.O.knn2(q, t, nthreads=threads)
        return len(q) * len(t) / (time.perf_counter() - t0)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import synth
    cpu = CpuReference()
    cores = host_cores()
    frames = synth.stereo_batch(2, W, H, BATCH // 2)
    for _ in range(args.warmup):
        cpu.extract_frames(frames[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.extract_frames(frames, cores)
    dt = time.perf_counter() - t0
    fps = BATCH * args.steps / dt
    q, t = synth.matching_set(NQ, NT)
    pairs = cpu.knn2(q[:200], t, cores)
    sample = f"{BATCH}-frame synthetic stereo batch per step on {cores} host threads, one extractor instance per thread"
    line = {"impl": "reference", "metric": "ORB frames/s (1241x376, 2000 kp)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "synthetic stereo pairs 1241x376, 2000 features, 8 levels, 64-frame batch (BASELINE config[1])",
                       "frames_per_step": BATCH},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu.kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matching": {"value": pairs, "unit": "pairs/s", "cores": cores, "sample": "200 of 2000 queries x 100000 train"}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(index):
    """Run this rank on the host cores next to its GPU (NVML's ideal CPU set) before any pinned buffer is allocated, so
    that first touch puts the staging memory on the GPU's own NUMA node: with 8 ranks the H2D streams otherwise cross the
    socket interconnect.  Returns the number of cores bound to, or None when NVML has no answer."""
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = N.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, v in enumerate(mask) for b in range(64) if (int(v) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import orbx
    import shard
    import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: liborbx has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: POOL distinct 64-frame stereo batches, pinned on the host and resident in HBM
    base = synth.stereo_batch(2 + rank, W, H, BATCH // 2)
    host_pool, dev_pool = [], []
    for p in range(POOL):
        hb = torch.empty((BATCH, H, W), dtype=torch.uint8).pin_memory()
        for f in range(BATCH):
            src = base[(f + 7 * p) % BATCH]
            hb[f] = torch.from_numpy(np.roll(src, 3 * p, axis=1) if p else src)
        host_pool.append(hb)
        dev_pool.append(hb.to(dev, non_blocking=True))
    torch.cuda.synchronize()

    ex = orbx.Extractor(NFEAT, 1.2, NLEVELS, 20, 7, max_width=W, max_height=H, max_batch=BATCH, device=local_rank)
    stream = torch.cuda.Stream(device=dev)

    def step_dev(i):
        d = dev_pool[i % POOL]
        ex.extract_batch_device(d.data_ptr(), H * W, W, BATCH, W, H, stream.cuda_stream)

    for i in range(args.warmup):
        step_dev(i)
    launches_per_step = ex.last_launches()   # counted by the library: copy + 7 resize + 2 FAST + blur + octree + describe per half batch
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)

    # ---- e2e through the host entry point (pinned host frames in, keypoints + descriptors out)
    cap = ex.max_keypoints
    # pinned result arrays with the library's stride: keypoints and descriptors are DMA'd straight into them
    out = (torch.zeros(BATCH * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(BATCH, cap),
           torch.zeros((BATCH, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(BATCH, np.int32))
    host_np = [[hb[f].numpy() for f in range(BATCH)] for hb in host_pool]
    host_ptrs = [orbx.Extractor.frame_pointers(fr) for fr in host_np]     # the `const uint8_t *const *` a C++ caller passes

    def step_e2e(i):
        ex.extract_batch_ptrs(host_ptrs[i % POOL], BATCH, W, H, W, out)

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    s_e2e = time.perf_counter() - t0
    n_kp = int(out[2].sum())
    launches_e2e = ex.last_launches()
    # the PCIe floor of the e2e number: one pinned H2D copy of a step's frames by itself
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dev_pool[0].copy_(host_pool[0], non_blocking=True)
        c0.record(stream)
        for p in range(5):
            dev_pool[p % POOL].copy_(host_pool[p % POOL], non_blocking=True)
        c1.record(stream)
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1) / 5

    clk = clocks.stop() if rank == 0 else None

    # ---- per-stage device times -> dominant kernel for the roofline line
    stages = ex.profile_stages(reps=5)

    # ---- latency of the call OrbFrame makes: one frame (and one stereo pair) through the host entry point
    lat = {}
    for nb, key in ((1, "single_frame_ms"), (2, "stereo_pair_ms")):
        ptrs = orbx.Extractor.frame_pointers(host_np[0][:nb])
        o = (out[0][:nb], out[1][:nb], out[2][:nb])
        for _ in range(10):
            ex.extract_batch_ptrs(ptrs, nb, W, H, W, o)
        t0 = time.perf_counter()
        for _ in range(100):
            ex.extract_batch_ptrs(ptrs, nb, W, H, W, o)
        lat[key] = (time.perf_counter() - t0) / 100 * 1e3

    # ---- matching: query-sharded kNN-2, one all_gather of the result records
    q, t = synth.matching_set(NQ, NT)
    m = orbx.Matcher(max_queries=NQ, max_train=NT, device=local_rank)
    nq_loc = shard.query_block(NQ, world)
    q_loc = np.zeros((nq_loc, 32), np.uint8)
    qlo, qhi = shard.query_range(NQ, rank, world)
    part = q[qlo:qhi]
    q_loc[:len(part)] = part
    dq = torch.from_numpy(q_loc).to(dev)
    dt_ = torch.from_numpy(t).to(dev)
    d_loc = torch.empty((nq_loc, 4), dtype=torch.int32, device=dev)
    d_all = torch.empty((world * nq_loc, 4), dtype=torch.int32, device=dev)

    def step_match():
        m.knn2_device(dq.data_ptr(), nq_loc, dt_.data_ptr(), NT, d_loc.data_ptr(), stream.cuda_stream)
        if world > 1:
            with torch.cuda.stream(stream):
                shard.gather_match_records(d_loc, NQ, out=d_all)

    for _ in range(max(args.warmup, 3)):
        step_match()
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    msteps = max(args.steps, 20)
    m0.record(stream)
    for _ in range(msteps):
        step_match()
    m1.record(stream)
    barrier()
    ms_match = m0.elapsed_time(m1)

    popc_rate = m.measure_popc()

    # ---- "next" rows, rank 0 only, reported beside the headline: bag-of-words descent over one frame's
    # descriptors (N2), distinctive descriptors of a map's points (N4), stereo matching of one pair (N1)
    next_rows = None
    if rank == 0:
        next_rows = {}
        rng = np.random.default_rng(99)
        voc = orbx.random_vocabulary(10, 5, seed=1)                       # k = 10, L = 5: 111 111 nodes (ORBvoc.txt is k = 10, L = 6)
        V = orbx.Vocabulary(*voc[:5], voc[5], device=local_rank)
        feats = rng.integers(0, 256, (BATCH * NFEAT, 32), dtype=np.uint8)
        dfe = torch.from_numpy(feats).to(dev)
        dwn = torch.empty((BATCH * NFEAT, 2), dtype=torch.int32, device=dev)
        for _ in range(3):
            V.transform_device(dfe.data_ptr(), 32, BATCH * NFEAT, 4, dwn.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(10):
            V.transform_device(dfe.data_ptr(), 32, BATCH * NFEAT, 4, dwn.data_ptr(), stream.cuda_stream)
        n1.record(stream)
        torch.cuda.synchronize()
        ms = n0.elapsed_time(n1) / 10
        next_rows["bow_transform"] = {"features_per_s": BATCH * NFEAT / (ms * 1e-3), "ms_per_64_frames": ms,
                                      "workload": f"{BATCH} frames x {NFEAT} descriptors, synthetic vocabulary k=10 L=5, levels_up=4, device-resident"}
        V.close()
        # stereo matching of the 32 L/R pairs of a batch (N1): pool 0 holds the pairs in L R L R order
        ex.extract_batch_ptrs(host_ptrs[0], BATCH, W, H, W, out)
        sout = (np.zeros((BATCH // 2, cap), np.float32), np.zeros((BATCH // 2, cap), np.float32))
        orbx.stereo_match_batch(ex, BATCH // 2, 0, 1, 2, 386.1448, 0.5372, out=sout)
        t0 = time.perf_counter()
        for _ in range(5):
            _, _, snl, snm = orbx.stereo_match_batch(ex, BATCH // 2, 0, 1, 2, 386.1448, 0.5372, out=sout)
        dt_s = (time.perf_counter() - t0) / 5
        next_rows["stereo_matches"] = {"pairs_per_s": (BATCH // 2) / dt_s, "ms_per_call": dt_s * 1e3, "matches_per_pair": float(snm.mean()),
                                       "workload": f"{BATCH // 2} stereo pairs of one extracted batch in one orbx_stereo_match_batch call, results to host"}
        npts = 20000
        pool = rng.integers(0, 256, (npts * 4, 32), dtype=np.uint8)
        sizes = rng.integers(2, 25, npts)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        inds = rng.integers(0, len(pool), int(offs[-1])).astype(np.int32)
        m.distinctive(pool, offs, inds)
        t0 = time.perf_counter()
        for _ in range(3):
            m.distinctive(pool, offs, inds)
        dt_d = (time.perf_counter() - t0) / 3
        next_rows["distinctive_descriptors"] = {"points_per_s": npts / dt_d, "ms_per_call": dt_d * 1e3,
                                                "workload": f"{npts} map points with 2..24 observations each, host arrays in and out"}

        # SearchByProjection (N3) with the frame grid on the device: frame 0's key points as map points, projected
        # 1.25 px beside themselves into frame 0 (the shape of TrackLocalMap: ~2000 points against ~2000 key points)
        k0 = np.ascontiguousarray(out[0][0, :out[2][0]]); d0 = np.ascontiguousarray(out[1][0, :out[2][0]])
        sfl = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
        proj_args = (k0, np.full(len(k0), -1, np.float32), None, d0, (0.0, 0.0, float(W), float(H)), d0, k0["x"] + np.float32(1.25), k0["y"].copy(),
                     k0["octave"].copy(), (np.float32(4.0) * sfl[k0["octave"]]).astype(np.float32))
        m.search_by_projection(*proj_args)
        t0 = time.perf_counter()
        for _ in range(20):
            _, _, pnm = m.search_by_projection(*proj_args)
        dt_p = (time.perf_counter() - t0) / 20
        next_rows["search_by_projection"] = {"map_points_per_s": len(k0) / dt_p, "ms_per_call": dt_p * 1e3, "matches": int(pnm),
                                             "workload": f"{len(k0)} map points against one frame of {len(k0)} key points, grid + candidate lists + best/second best on the device, host arrays in and out"}

    # ---- max over ranks
    times = torch.tensor([ms_dev, s_e2e, ms_match], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, s_e2e, ms_match = [float(x) for x in times.tolist()]

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        frames = BATCH * args.steps * world
        fps_dev = frames / (ms_dev * 1e-3)
        fps_e2e = frames / s_e2e
        dom = max(stages, key=stages.get)
        dom_ms = stages[dom]
        ach = B_ALG * BATCH / (dom_ms * 1e-3) / 1e9
        traffic, winst_step = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            prof = json.load(open(tp))
            traffic = prof.get("bytes_per_step", {}).get(dom)
            winst_step = sum(prof.get("warp_instructions_per_step", {}).values()) or None
        pairs = NQ * NT * msteps / (ms_match * 1e-3)
        int_peak = 148 * sm_max * 1e6 * popc_rate / 8      # pairs/s at the POPC rate measured on this GPU, SURVEY 8(d)
        line = {
            "metric": "ORB frames/s (1241x376, 2000 kp)", "value": fps_dev, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "synthetic stereo pairs 1241x376, 2000 features, 8 levels, 64-frame batch per GPU per step (BASELINE config[1])",
                       "frames_per_gpu_per_step": BATCH, "global_frames_per_step": BATCH * world,
                       "l2": f"inputs rotate over {POOL} resident batches ({POOL * BATCH * W * H / 1e6:.0f} MB) + 2x{BATCH * 2.2:.0f} MB of pyramid/blur slabs rewritten every step: working set > 126 MB L2",
                       "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                       "host_binding": None if numa is None else f"each rank bound to the {numa} host cores NVML lists for its GPU"},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": BATCH * W * H,
                    "d2h_bytes_per_step": BATCH * cap * 60 + BATCH * 4, "keypoints_last_step": n_kp,
                    "h2d_alone_ms_per_step": h2d_ms, "h2d_gbs": BATCH * W * H / (h2d_ms * 1e-3) / 1e9,
                    "pcie_floor_frames_per_s": BATCH * world / (h2d_ms * 1e-3)},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": {"device_resident": launches_per_step, "e2e": launches_e2e},
            "latency": dict(lat, note="orbx_extract_batch with 1 / 2 frames of 1241x376, pinned host buffers, H2D + kernels + D2H, mean of 100 calls"),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full of this stage, bytes per 64-frame step)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B_ALG * BATCH, "kernel_ms": dom_ms,
                         "whole_step_frac": B_ALG * BATCH / (ms_dev / args.steps * 1e-3) / 1e9 / peak,
                         # the path is instruction-issue bound (50-100 integer operations per byte): warp instructions of one
                         # step (ncu smsp__inst_executed.sum, profiles/ncu_traffic.json) against 4 issue slots x 148 SMs x clock
                         "issue_roofline": None if not winst_step else {
                             "warp_instructions_per_step": winst_step,
                             "peak_warp_instructions_per_s": 4 * 148 * sm_max * 1e6,
                             "frac": winst_step / (ms_dev / args.steps * 1e-3) / (4 * 148 * sm_max * 1e6)}},
            "stages_ms": stages,
            "clocks": clk,
            "matching": {"metric": "Hamming kNN-2 pairs/s (2000 x 100000)", "value": pairs, "unit": "pairs/s",
                         "queries_per_s": NQ * msteps / (ms_match * 1e-3), "ms_per_batch": ms_match / msteps,
                         "scaling": "strong", "sharding": f"{nq_loc} queries per GPU, train replicated, all_gather of 16-byte records" if world > 1 else "single GPU",
                         "roofline": {"bound": "int", "achieved": pairs / world, "peak": int_peak, "unit": "pairs/s/GPU",
                                      "frac": pairs / world / int_peak, "popc_per_clk_per_sm": popc_rate,
                                      "peak_source": "148 SM x max SM clock x POPC lanes/clk/SM measured on this GPU (orbm_measure_popc) / 8 POPC per pair"}},
        }
        line["next_rows"] = next_rows
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuReference()
            cores = host_cores()
            frames_cpu = base * 4                      # 256 frames
            fps_all = cpu.extract_frames(frames_cpu, cores)
            fps_one = cpu.extract_frames(base[:16], 1)
            line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": cpu.kind,
                                    "sample": f"{len(frames_cpu)} frames of the same workload over {cores} host threads (one extractor instance per thread); single thread: {fps_one:.1f} frames/s on 16 frames",
                                    "single_thread": fps_one,
                                    "matching_pairs_per_s": cpu.knn2(q[:160], t, cores)}
            # the same next-row workloads on one host core through the oracle's restatements (bounded samples)
            O = cpu.O
            voc = orbx.random_vocabulary(10, 5, seed=1)
            fs = np.random.default_rng(99).integers(0, 256, (20000, 32), dtype=np.uint8)
            t0 = time.perf_counter(); O.voc_transform(voc[0], voc[1], voc[2], voc[3], voc[5], 4, fs); dtv = time.perf_counter() - t0
            line["next_rows"]["bow_transform"]["cpu_port_features_per_s_1core"] = len(fs) / dtv
            t0 = time.perf_counter(); O.distinctive(pool, offs[:2001], inds[:offs[2000]]); dtd = time.perf_counter() - t0
            line["next_rows"]["distinctive_descriptors"]["cpu_port_points_per_s_1core"] = 2000 / dtd
            t0 = time.perf_counter()
            for _ in range(5):
                O.search_by_projection(*proj_args)
            line["next_rows"]["search_by_projection"]["cpu_port_map_points_per_s_1core"] = len(k0) * 5 / (time.perf_counter() - t0)
        print(json.dumps(line), flush=True)
    ex.close(); m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
